#!/usr/bin/env python
"""bench.py -- the contract benchmark: pairs sorted per second (64-bit key + 64-bit rid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one complete sort of one batch of synthetic pairs.
  N = 1   workload = BASELINE.json configs[1]: 2^30 uniform 64-bit key+rid pairs on one B200.
  N > 1   (torchrun, one rank per GPU) every rank holds --pairs-per-gpu pairs (weak scaling);
          the pairs are range-partitioned over NCCL and every rank sorts its range.
`value`   device-resident throughput (inputs in HBM when the timed region starts), CUDA events.
`e2e`     the same sort through the reference-facing C-ABI call sort() with pinned HOST
          arrays: host->device and device->host copies inside the timed region.
`roofline`    the dominant kernel (scatter): algorithmic bytes / CUDA-event time, against
              MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the unmodified reference (oracle/_ref, msb_64.c with its
              64 threads) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pairs_sorted_per_second_64bit_key_rid"
UNIT = "Gpairs/s"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, "of fallback"


def parse_count(text) -> int:
    """'1<<30', '2**30', '3*(1<<28)', '1073741824' ... -> int, without eval()."""
    import ast
    import operator as op
    ops = {ast.LShift: op.lshift, ast.Mult: op.mul, ast.Pow: op.pow, ast.Add: op.add, ast.Sub: op.sub,
           ast.FloorDiv: op.floordiv}

    def ev(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, int):
            return node.value
        if isinstance(node, ast.BinOp) and type(node.op) in ops:
            return ops[type(node.op)](ev(node.left), ev(node.right))
        raise ValueError(f"not a pair count: {text!r}")
    return int(ev(ast.parse(str(text), mode="eval").body))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=str, default="1<<30")
    ap.add_argument("--cpu-sample", type=str, default="1<<27",
                    help="pairs per step of the CPU reference (it refuses < 2^25)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: fused peer-memory route+exchange kernel, or route + NCCL all-to-all")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-child", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, gpu_index: int = 0) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9 or not p[0].isdigit() or int(p[0]) != gpu_index:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax = float(p[2])
                except ValueError:
                    continue
                for name, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        # samples under load only (idle samples before/after would drag the median down)
        loaded = [x for x in sm if smax and x > 0.5 * smax] or sm
        med = loaded[len(loaded) // 2] if loaded else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, "fallback"


def recorded_traffic():
    """dram bytes per scatter launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "scatter_traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------ CPU reference arm
class ReferenceRunner:
    """The unmodified reference on n uniform pairs per step (buffers allocated once)."""

    def __init__(self, n):
        import numpy as np
        from oracle import oracle as orc
        self.np, self.n = np, n
        self.ref = orc.RefLib()
        self.fudge = max(1.5, orc.min_fudge(n) + 0.05)
        cap = int(n * self.fudge) + 8192
        self.keys, self.rids = self.ref.aligned(cap), self.ref.aligned(cap)
        self.rng = np.random.default_rng(2026)

    def step(self) -> float:
        np, n = self.np, self.n
        self.keys[:n] = self.rng.integers(0, 1 << 64, size=n, dtype=np.uint64)
        self.rids[:n] = np.arange(n, dtype=np.uint64)
        expect = int(np.sum(self.keys[:n], dtype=np.uint64))
        t0 = time.perf_counter()
        size, _ = self.ref.sort(self.keys, self.rids, n, self.fudge, threads=64)
        dt = time.perf_counter() - t0
        # acceptance as in the reference's own check(): ascending + key checksum
        k = self.keys[:n]
        assert size == n and bool(np.all(k[:-1] <= k[1:]))
        assert int(np.sum(k, dtype=np.uint64)) == expect
        return dt


def ref_child(sample_n: int, steps: int) -> int:
    """Hidden mode (--ref-child): the reference in a process of its own, one JSON line per step.
    The unmodified msb_64.c draws its sample with a seed it never initialises
    (thread_data_t.seed) and, with its 64 threads oversubscribed on a small host, has been seen
    to return a misordered pair or to crash once in a few dozen runs -- so it is kept out of
    the process that holds the GPU results, every step is checked, and the parent retries."""
    runner = ReferenceRunner(sample_n)
    for _ in range(steps):
        try:
            dt = runner.step()
            print(json.dumps({"ok": True, "seconds": dt}), flush=True)
        except AssertionError:
            print(json.dumps({"ok": False}), flush=True)
    return 0


def run_reference_steps(sample_n: int, steps: int):
    """`steps` validated reference sorts of sample_n pairs: (seconds per good step, failures)."""
    good, failed, launches = [], 0, 0
    while len(good) < steps and launches < steps + 4:
        launches += 1
        want = steps - len(good)
        try:
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-child",
                                  "--cpu-sample", str(sample_n), "--steps", str(want)],
                                 capture_output=True, text=True, timeout=600)
            out = res.stdout
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        seen = 0
        for line in out.splitlines():
            try:
                rec = json.loads(line)
            except ValueError:
                continue
            seen += 1
            if rec.get("ok"):
                good.append(float(rec["seconds"]))
            else:
                failed += 1
        if seen < want:
            failed += 1                                  # the child died inside a step
    return good, failed


def cpu_baseline(sample_n: int, steps: int = 1, warmup: int = 0):
    """The reference msb_64 on the host cores; falls back to the oracle port."""
    import numpy as np
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(2026)
    if os.path.exists(orc.REF_SO):
        times, failed = run_reference_steps(sample_n, steps + warmup)
        times = times[warmup:] if len(times) > warmup else times
        if times:
            dt = sum(times) / len(times)
            note = (f"; {failed} further run(s) of the reference failed its own check() or crashed "
                    f"and were repeated" if failed else "")
            return {"value": sample_n / dt / 1e9, "unit": UNIT, "cores": min(64, cores),
                    "kind": "reference",
                    "sample": f"{sample_n} uniform pairs per step, oracle/_ref (unmodified msb_64.c, "
                              f"64 threads on {cores} host cores), mean of {len(times)} checked step(s)"
                              + note,
                    "seconds_per_step": dt}
    o = orc.Oracle()
    n = min(sample_n, 1 << 22)
    keys = rng.integers(0, 1 << 64, size=n + n // 2 + 64, dtype=np.uint64)
    rids = np.arange(keys.size, dtype=np.uint64)
    t0 = time.perf_counter()
    o.sort([keys], [rids], [n])
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} uniform pairs, oracle/msb64_oracle.c single thread"
                      + (" (the compiled reference kept failing)" if os.path.exists(orc.REF_SO) else ""),
            "seconds_per_step": dt}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_n = parse_count(args.cpu_sample)
    per_gpu = parse_count(args.pairs_per_gpu)
    base = cpu_baseline(sample_n, steps=max(args.steps, 1), warmup=max(min(args.warmup, 1), 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"uniform 64-bit key+rid pairs, {per_gpu} per GPU "
                               f"(BASELINE.json configs[1]); CPU arm sorts a bounded sample "
                               f"of {sample_n} pairs per step"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ B200 arm
def main_b200(args):
    import torch
    import inplacemsdradixsort_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    if not torch.cuda.is_available() or m.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = parse_count(args.pairs_per_gpu)
    stream = torch.cuda.current_stream().cuda_stream
    lib = m.load_library()

    def fill(kt, rt, seed):
        rc = lib.msb64_b200_fill(kt.data_ptr(), rt.data_ptr(), kt.numel(), 0, seed, 0, stream)
        assert rc == 0, lib.msb64_b200_last_error()

    src_k = torch.empty(n, dtype=torch.int64, device=dev)
    src_r = torch.empty(n, dtype=torch.int64, device=dev)
    fill(src_k, src_r, 1000 + rank)
    if dist is not None:
        src_r += rank * n                                   # globally unique rids
    keys = torch.empty_like(src_k)
    rids = torch.empty_like(src_r)

    def check(kt, rt, cnt):
        out = (ctypes.c_uint64 * 3)()
        rc = lib.msb64_b200_check(kt.data_ptr(), rt.data_ptr(), cnt, out, stream)
        assert rc == 0, lib.msb64_b200_last_error()
        return int(out[0]), int(out[1]), int(out[2])

    _, sum0, dig0 = check(src_k, src_r, n)

    sorter = None
    if dist is not None:
        from inplacemsdradixsort_b200.distributed import ShardedSorter
        sorter = ShardedSorter(n, dev, exchange=args.exchange)

    def one_step():
        if sorter is None:
            m.sort_tensors(keys, rids)
            return keys, rids, n
        return sorter.sort(keys, rids)

    def restore():
        keys.copy_(src_k)
        rids.copy_(src_r)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        restore()
        one_step()
    barrier()

    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    launches0 = m.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        restore()                       # outside the per-step events: not part of the sort
        if dist is not None:
            dist.barrier()
        a.record()
        out_k, out_r, out_n = one_step()
        b.record()
    barrier()
    launches = m.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())

    # correctness of the last timed step: ascending, checksum, (key, rid) multiset digest
    bad, sum1, dig1 = check(out_k, out_r, out_n)
    ok = bad == 0
    if dist is None:
        ok = ok and sum1 == sum0 and dig1 == dig0
    else:
        def wrap_sum(x):
            t = torch.tensor([x & 0xFFFFFFFF, x >> 32], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            lo, hi = int(t[0]), int(t[1])
            return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF
        ok_all = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        ok = bool(ok_all.item()) and wrap_sum(sum1) == wrap_sum(sum0) \
            and wrap_sum(dig1) == wrap_sum(dig0)
        ok = ok and sorter.boundaries_ordered(out_k, out_n)
    if not ok:
        raise SystemExit(f"rank {rank}: sorted output failed verification")

    clocks = sampler.stop(local_rank) if rank == 0 else None
    value = world * n * args.steps / (total_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": f"uniform 64-bit key + 64-bit rid pairs, {n} per GPU " + (
                "(BASELINE.json configs[4]: 2^34 pairs over 8 GPUs)" if n * world == 1 << 34 and world == 8 else
                f"(BASELINE.json configs[1]{'' if world == 1 else ', one such shard per GPU x' + str(world)})"),
            "pairs_per_gpu": n, "total_pairs": n * world,
            "l2": "inputs (16 B x pairs per GPU) far exceed the 126 MB L2; every step re-reads "
                  "a fresh unsorted copy",
            "schedule_bits": m.get_schedule(n),
            "parallelism": "single GPU" if world == 1 else
                           f"range partition over {world} GPUs, exchange = " +
                           ("routing kernel storing into the peers' HBM over NVLink (CUDA IPC)"
                            if sorter.exchange == "peer" else "route kernel + NCCL all-to-all"),
        },
        "clocks": clocks, "gpu_launches": launches, "verified": True,
    }

    # ---- N > 1: device times of the sharded steps; the exchange against NVLink bandwidth
    if sorter is not None:
        restore()
        barrier()
        sorter.sort(keys, rids, timed=True)
        lt = sorter.last_times
        tt = torch.tensor([lt["plan"], lt["exchange"], lt["barrier"], lt["local_sort"]],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sent = torch.tensor([lt["pairs_sent_to_peers"]], dtype=torch.int64, device=dev)
        dist.all_reduce(sent, op=dist.ReduceOp.MAX)
        ex_ms = float(tt[1] + tt[2])
        gbs = 16 * int(sent.item()) / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None
        line["exchange"] = {
            "kind": sorter.exchange, "plan_ms": float(tt[0]), "exchange_ms": float(tt[1]),
            "barrier_ms": float(tt[2]), "local_sort_ms": float(tt[3]),
            "bytes_out_per_gpu": 16 * int(sent.item()), "out_GB/s_per_gpu": gbs,
            "nvlink_peak_GB/s_per_direction": 900.0,
            "frac_of_nvlink": gbs / 900.0 if gbs else None,
            "note": "max over ranks; exchange = routing kernel (+ all-to-all for nccl) + completion barrier",
        }

    # ---- roofline of the dominant kernel, from CUDA events inside this process
    if rank == 0 or dist is None:
        restore()
        torch.cuda.synchronize()
        phases = m.sort_device(keys.data_ptr(), rids.data_ptr(), n, stream=stream, timed=True)
        levels = m.last_level_times()
        stats = m.last_stats()
        peak, how = measured_hbm_peak()
        moved = stats.get("moved", [])
        active = [(moved[l], levels[l]["scatter"]) for l in range(len(levels)) if moved[l]]
        kernels = {}
        if active:
            byts = sum(32 * p for p, _ in active)
            us = sum(t for _, t in active)
            ach = byts / (us * 1e-6) / 1e9
            traffic = recorded_traffic()
            line["roofline"] = {
                "bound": "hbm", "kernel": "scatter_kernel", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic else None,
                "peak_source": f"{how} (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "launches": len(active), "avg_launch_ms": us / len(active) / 1e3,
                "algorithmic_bytes_per_launch": byts / len(active),
            }
            kernels["scatter"] = {"GB/s": ach, "frac": ach / peak, "ms": us / 1e3}
            hus = sum(levels[l]["histogram"] for l in range(len(levels)) if moved[l])
            # keys the histogram passes really read (level 1 is counted inside level 0's pass)
            hb = 8 * stats["hist_keys"] if stats.get("hist_keys") else sum(8 * p for p, _ in active)
            if hus:
                kernels["histogram"] = {"GB/s": hb / (hus * 1e-6) / 1e9,
                                        "frac": hb / (hus * 1e-6) / 1e9 / peak, "ms": hus / 1e3}
        if phases["local_sort"] and stats.get("local_pairs"):
            lb = 32 * stats["local_pairs"]
            g = lb / (phases["local_sort"] * 1e-6) / 1e9
            kernels["local_sort"] = {"GB/s": g, "frac": g / peak, "ms": phases["local_sort"] / 1e3}
        kernels["plan"] = {"ms": phases["plan"] / 1e3}
        line["kernels"] = kernels
        line["phases_us"] = phases

    # ---- end to end through the reference-facing C ABI with host buffers
    if not args.no_e2e:
        hk, hr = m.pinned(n), m.pinned(n)
        e2e_s = []
        for i in range(args.e2e_steps + 1):
            rc = lib.msb64_b200_memcpy_d2h(hk.ctypes.data, src_k.data_ptr(), n * 8, stream)
            rc |= lib.msb64_b200_memcpy_d2h(hr.ctypes.data, src_r.data_ptr(), n * 8, stream)
            assert rc == 0
            barrier()
            t0 = time.perf_counter()
            size = [n]
            m.sort([hk], [hr], size)             # H2D + sort + D2H, synchronous
            dt = time.perf_counter() - t0
            if i:                                # first call is warm-up (allocations)
                e2e_s.append(dt)
        import numpy as np
        assert bool(np.all(hk[:-1] <= hk[1:])), "e2e output not sorted"
        t = torch.tensor([sum(e2e_s) / len(e2e_s)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": world * n / float(t.item()) / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": 16 * n * world,
                       "ms_per_step": float(t.item()) * 1e3, "steps": len(e2e_s),
                       "api": "sort() of include/msb64_b200.h with pinned host arrays"
                              + ("" if world == 1 else " (each rank sorts its own shard; no exchange)")}
        m.free_pinned(hk)
        m.free_pinned(hr)

    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(parse_count(args.cpu_sample)).items()
                                    if k != "seconds_per_step"}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(),
                                    "kind": "reference", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    if a.ref_child:
        sys.exit(ref_child(parse_count(a.cpu_sample), a.steps))
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
