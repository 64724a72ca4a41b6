#!/usr/bin/env python
"""bench.py -- the contract benchmark: pairs sorted per second (64-bit key + 64-bit rid).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step is one complete sort of one batch of synthetic pairs.
  N = 1   workload = BASELINE.json configs[1]: 2^30 uniform 64-bit key+rid pairs on one B200.
          The default line also carries `workloads`: the other inputs BASELINE.json names
          (configs[2], [3]: few distinct values, heavy duplicates, low 24 bits, presorted,
          reverse sorted) and three adversarial ones, each timed and verified at the same size.
  N > 1   (torchrun, one rank per GPU) every rank holds --pairs-per-gpu pairs (weak scaling;
          --pairs-per-gpu '1<<31' with --gpus 8 is BASELINE.json configs[4], 2^34 pairs);
          the pairs are range-partitioned across the GPUs and every rank sorts its range.
`value`   device-resident throughput (inputs in HBM when the timed region starts), CUDA events.
`e2e`     N = 1: the same sort through the reference-facing C-ABI call sort() with pinned HOST
          arrays; N > 1: pinned host arrays -> the sharded sort -> pinned host arrays on every
          rank.  Host->device and device->host copies inside the timed region.
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

# name -> (msb64_b200_fill kind, param, what it is / which BASELINE.json config it covers)
WORKLOADS = {
    "uniform": (0, 0, "uniform 64-bit keys (configs[1])"),
    "dup16": (2, 16, "16 distinct values (configs[2]: few distinct values)"),
    "dup1e6": (2, 10 ** 6, "10^6 distinct values (configs[2]: heavy duplicates)"),
    "low24": (1, (1 << 24) - 1, "only the low 24 bits significant (configs[3])"),
    "sorted": (3, 1, "presorted, key = index (configs[3])"),
    "reverse": (4, 1, "reverse sorted (configs[3])"),
    "clustered": (5, 0, "clusters of ~3000 keys differing in their low 2 bits (adversarial for the local sort)"),
    "outlier": (6, 0, "12-bit keys and one outlier at 2^63 (adversarial for the digit positions)"),
    "zipf": (7, 0, "zipf-like, exponent 1.3 (skewed bucket sizes)"),
}


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
    ap.add_argument("--cpu-sample", type=str, default="1<<28",
                    help="pairs per step of the CPU reference (BASELINE.json configs[0]: 2^28; it refuses < 2^25)")
    ap.add_argument("--workload", default="uniform", choices=sorted(WORKLOADS),
                    help="input of the timed steps (the other ones ride along in `workloads` at N = 1)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the `workloads` object")
    ap.add_argument("--exchange", default="auto", choices=["auto", "pipelined", "peer", "nccl"],
                    help="N > 1: bucket pass + copy-engine exchange overlapped with the sorts (pipelined), "
                         "fused peer-store route kernel (peer), or route + NCCL all-to-all (nccl)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-child", action="store_true", help=argparse.SUPPRESS)
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under the bench's load.

    The sampler is started in front of the warm-up steps (nvidia-smi needs a few hundred
    milliseconds before its first sample, more than a short timed region lasts) and stopped
    behind the timed region; mark() notes when the timed region begins, and the samples from
    then on are the ones reported when there are any -- otherwise the samples of warm-up +
    timed region together (same kernels, same load), which `window` says."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc = None
        self.path = None
        self.marked = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def mark(self):
        self.marked = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, gpu_index: int = 0) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows, smax = [], None
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 10 or not p[1].isdigit() or int(p[1]) != gpu_index:
                    continue
                try:
                    mhz, smax = float(p[2]), float(p[3])
                except ValueError:
                    continue
                rows.append((self._stamp(p[0]), mhz,
                             {n for n, v in zip(names, p[6:10]) if v.lower().startswith("active")}))
        os.unlink(self.path)
        timed = [r for r in rows if self.marked and r[0] and r[0] >= self.marked - 0.02]
        window = "timed region"
        use = [r for r in timed if smax and r[1] > 0.5 * smax]
        if len(use) < 3:
            # too short a region for nvidia-smi's sampling: warm-up steps included
            use = [r for r in rows if smax and r[1] > 0.5 * smax] or rows
            window = "warm-up + timed region (the timed region alone gave fewer than 3 samples)"
        sm = sorted(r[1] for r in use)
        reasons = set().union(*[r[2] for r in use]) if use else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(use), "window": window}


def near_cpus(gpu_index: int):
    """CPUs NVML names as closest to the GPU (intersected with what this process may use), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


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
    """The unmodified reference on n uniform pairs per step (buffers allocated once).  Keys come
    from the reference's own generator (rand.c, MT19937-64; restated in oracle/msb64_oracle.c
    and pinned against the compiled rand.c by tests/test_oracle_pin.py), a new seed per step --
    BASELINE.json configs[0]'s input."""

    def __init__(self, n):
        import numpy as np
        from oracle import oracle as orc
        self.np, self.n = np, n
        self.ref = orc.RefLib()
        self.gen = orc.Oracle()
        self.fudge = max(1.5, orc.min_fudge(n) + 0.05)
        cap = int(n * self.fudge) + 8192
        self.keys, self.rids = self.ref.aligned(cap), self.ref.aligned(cap)
        self.seed = 2026

    def step(self) -> float:
        np, n = self.np, self.n
        self.seed += 1
        self.keys[:n] = self.gen.rand64(self.seed, n)
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
                                 capture_output=True, text=True, timeout=1200)
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
    if os.path.exists(orc.REF_SO):
        times, failed = run_reference_steps(sample_n, steps + warmup)
        times = times[warmup:] if len(times) > warmup else times
        if times:
            dt = sum(times) / len(times)
            return {"value": sample_n / dt / 1e9, "unit": UNIT, "cores": min(64, cores),
                    "kind": "reference",
                    "sample": f"{sample_n} uniform pairs per step from the reference's rand.c generator, "
                              f"oracle/_ref (unmodified msb_64.c, 64 threads on {cores} host cores), "
                              f"mean of {len(times)} checked step(s)",
                    "reference_retries": failed,
                    "seconds_per_step": dt}
    o = orc.Oracle()
    n = min(sample_n, 1 << 22)
    keys = np.concatenate([o.rand64(2026, n), np.zeros(n // 2 + 64, dtype=np.uint64)])
    rids = np.arange(keys.size, dtype=np.uint64)
    t0 = time.perf_counter()
    o.sort([keys], [rids], [n])
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} uniform pairs (rand.c generator), oracle/msb64_oracle.c single thread"
                      + (" (the compiled reference kept failing)" if os.path.exists(orc.REF_SO) else ""),
            "reference_retries": 0,
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
        "config": {"workload": f"uniform 64-bit key + 64-bit rid pairs, {per_gpu} per GPU "
                               f"(BASELINE.json configs[1]); the CPU arm sorts {sample_n} pairs per step"
                               + (" = BASELINE.json configs[0] (2^28 pairs, rand.c generator, 64 threads)"
                                  if sample_n == 1 << 28 else " (a bounded sample)"),
                   "cpu_pairs_per_step": sample_n},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "reference_retries": base["reference_retries"],
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ B200 arm
def main_b200(args):
    # stdout carries the one JSON line and nothing else: whatever libraries print while they
    # initialise (NCCL announces its version on stdout) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
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

    def fill(kt, rt, seed, workload="uniform"):
        kind, param, _ = WORKLOADS[workload]
        rc = lib.msb64_b200_fill(kt.data_ptr(), rt.data_ptr(), kt.numel(), kind, seed, param, stream)
        assert rc == 0, lib.msb64_b200_last_error()

    src_k = torch.empty(n, dtype=torch.int64, device=dev)
    src_r = torch.empty(n, dtype=torch.int64, device=dev)
    fill(src_k, src_r, 1000 + rank, args.workload)
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

    def wrap_sum(x):
        t = torch.tensor([x & 0xFFFFFFFF, x >> 32], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        lo, hi = int(t[0]), int(t[1])
        return (lo + (hi << 32)) & 0xFFFFFFFFFFFFFFFF

    def verified(out_k, out_r, out_n, want_sum, want_dig):
        """ascending, key checksum, (key, rid) multiset digest; across ranks: boundaries ordered"""
        bad, sum1, dig1 = check(out_k, out_r, out_n)
        ok = bad == 0
        if dist is None:
            return ok and sum1 == want_sum and dig1 == want_dig
        ok_all = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        ok = bool(ok_all.item()) and wrap_sum(sum1) == wrap_sum(want_sum) and wrap_sum(dig1) == wrap_sum(want_dig)
        return ok and sorter.boundaries_ordered(out_k, out_n)

    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 0)):
        restore()
        one_step()
    barrier()
    sampler.mark()
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

    # correctness of the last timed step
    if not verified(out_k, out_r, out_n, sum0, dig0):
        raise SystemExit(f"rank {rank}: sorted output failed verification")
    if lib.msb64_b200_last_status(stream) != 0:
        raise SystemExit(f"rank {rank}: {lib.msb64_b200_last_error().decode()}")

    clocks = sampler.stop(local_rank) if rank == 0 else None
    value = world * n * args.steps / (total_ms * 1e-3) / 1e9
    what = WORKLOADS[args.workload][2]
    if args.workload == "uniform":
        cfg = ("(BASELINE.json configs[4]: 2^34 pairs over 8 GPUs)" if n * world == 1 << 34 and world == 8 else
               f"(BASELINE.json configs[1]{'' if world == 1 else ', one such shard per GPU x' + str(world)})")
    else:
        cfg = f"({what})"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: 64-bit key + 64-bit rid pairs, {n} per GPU {cfg}",
            "pairs_per_gpu": n, "total_pairs": n * world,
            "l2": "inputs (16 B x pairs per GPU) far exceed the 126 MB L2; every step re-reads "
                  "a fresh unsorted copy",
            "schedule_bits": m.get_schedule(n),
            "parallelism": "single GPU" if world == 1 else
                           f"range partition over {world} GPUs, exchange = " + {
                               "pipelined": "local bucket pass, then copy engines move the buckets into the peers' "
                                            "HBM over NVLink (CUDA IPC) while arrived sub-ranges are sorted",
                               "peer": "routing kernel storing into the peers' HBM over NVLink (CUDA IPC)",
                               "nccl": "route kernel + NCCL all-to-all"}[sorter.exchange],
        },
        "clocks": clocks, "gpu_launches": launches, "verified": True,
    }

    # ---- N > 1: device times of the sharded steps; the exchange against NVLink bandwidth
    if sorter is not None:
        restore()
        barrier()
        sorter.sort(keys, rids, timed=True)
        lt = sorter.last_times
        sent = torch.tensor([lt["pairs_sent_to_peers"] or 0], dtype=torch.int64, device=dev)
        dist.all_reduce(sent, op=dist.ReduceOp.MAX)
        out_bytes = 16 * int(sent.item())
        if sorter.exchange == "pipelined":
            names = ["plan", "route", "first_wait", "sort", "exchange", "step_device", "total"]
            tt = torch.tensor([lt[k] for k in names], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = dict(zip(names, (float(x) for x in tt)))
            direct = torch.tensor([lt.get("pairs_stored_by_route") or 0], dtype=torch.int64, device=dev)
            dist.all_reduce(direct, op=dist.ReduceOp.MAX)
            direct_bytes = 16 * int(direct.item())
            staged_bytes = out_bytes - direct_bytes
            gbs = staged_bytes / (t["exchange"] * 1e-3) / 1e9 if t["exchange"] > 0 else None
            # everything has left the GPU when the last copy is done: route pass + copies
            whole = out_bytes / ((t["route"] + t["exchange"]) * 1e-3) / 1e9 if t["exchange"] > 0 else None
            line["exchange"] = {
                "kind": "pipelined", "plan_ms": t["plan"], "route_ms": t["route"],
                "exchange_ms": t["exchange"], "first_wait_ms": t["first_wait"], "sort_ms": t["sort"],
                "step_ms": t["total"],
                "exchange_hidden_ms": max(t["exchange"] - t["first_wait"], 0.0),
                "exchange_hidden_frac": max(t["exchange"] - t["first_wait"], 0.0) / t["exchange"] if t["exchange"] > 0 else None,
                "bytes_out_per_gpu": out_bytes,
                "bytes_stored_by_route_kernel": direct_bytes,
                "route_kernel_out_GB/s": direct_bytes / (t["route"] * 1e-3) / 1e9 if t["route"] > 0 else None,
                "bytes_moved_by_copy_engines": staged_bytes, "copy_engines_out_GB/s": gbs,
                "out_GB/s_per_gpu": whole,
                "nvlink_peak_GB/s_per_direction": 900.0,
                "frac_of_nvlink": whole / 900.0 if whole else None,
                "note": "max over ranks, one timed step; plan = histogram + all-gather + host cut; route = bucket "
                        "pass, which stores the first quarter of every peer's sub-ranges straight into the peer's "
                        "HBM over NVLink; exchange = first to last outgoing copy of the staged rest (copy engines, "
                        "NVLink); out_GB/s_per_gpu = all outgoing bytes over route + exchange time; first_wait = main "
                        "stream idle until sub-range 0 is complete; sort = the sub-range sorts (32 per GPU up to 4 "
                        "GPUs, 16 on 8), running while the later sub-ranges are still travelling "
                        "(exchange_hidden_ms of the exchange)",
            }
        else:
            tt = torch.tensor([lt["plan"], lt["exchange"], lt["barrier"], lt["local_sort"]],
                              dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ex_ms = float(tt[1] + tt[2])
            gbs = out_bytes / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None
            line["exchange"] = {
                "kind": sorter.exchange, "plan_ms": float(tt[0]), "exchange_ms": float(tt[1]),
                "barrier_ms": float(tt[2]), "local_sort_ms": float(tt[3]),
                "bytes_out_per_gpu": out_bytes, "out_GB/s_per_gpu": gbs,
                "nvlink_peak_GB/s_per_direction": 900.0,
                "frac_of_nvlink": gbs / 900.0 if gbs else None,
                "note": "max over ranks; exchange = routing kernel (+ all-to-all for nccl) + completion barrier",
            }

    # ---- roofline of the dominant kernel, from CUDA events inside this process
    # (a single-GPU sort of one rank's share needs its own workspace: skipped where the sharded
    # sort's buffers leave no room for it, e.g. 2^31 pairs per GPU)
    room = torch.cuda.mem_get_info(dev)[0] > m.workspace_bytes(n) + (2 << 30) or dist is None
    if (rank == 0 or dist is None) and not room:
        line["roofline"] = None
        line["kernels"] = {"note": "not re-measured: no room for a single-GPU workspace beside the sharded sort's buffers"}
    if (rank == 0 or dist is None) and room:
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
            same = bool(traffic) and traffic.get("pairs") == n and args.workload == "uniform"
            line["roofline"] = {
                "bound": "hbm", "kernel": "scatter_kernel", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak,
                # dram bytes per launch of the committed ncu --set full capture; only quoted when
                # that capture was taken on this workload and size (it is not re-measured here;
                # profiles/README.md says on which build it was taken)
                "traffic": traffic.get("dram_bytes_per_launch") if same else None,
                "traffic_source": (traffic.get("source") if same else
                                   "none for this workload / size (profiles/scatter_traffic.json holds 2^30 uniform)"),
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

    # ---- the other inputs BASELINE.json names, same size, device-resident, verified (N = 1)
    if dist is None and not args.no_workloads and args.workload == "uniform":
        table = {}
        for name in WORKLOADS:
            if name == "uniform":
                table[name] = {"ms": total_ms / args.steps, "Gpairs/s": value, "verified": True,
                               "what": WORKLOADS[name][2]}
                continue
            fill(src_k, src_r, 77, name)
            _, s0, d0 = check(src_k, src_r, n)
            ms = []
            for i in range(3):
                restore()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                one_step()
                b.record()
                torch.cuda.synchronize()
                if i:
                    ms.append(a.elapsed_time(b))
            okw = verified(keys, rids, n, s0, d0) and lib.msb64_b200_last_status(stream) == 0
            table[name] = {"ms": sum(ms) / len(ms), "Gpairs/s": n / (sum(ms) / len(ms) * 1e-3) / 1e9,
                           "verified": bool(okw), "what": WORKLOADS[name][2]}
            if not okw:
                raise SystemExit(f"workload {name}: sorted output failed verification")
        line["workloads"] = table
        fill(src_k, src_r, 1000 + rank, args.workload)

    # ---- end to end with host buffers
    if not args.no_e2e:
        cap = n if sorter is None else sorter.recv_cap
        # a rank's host arrays belong on the NUMA node its GPU hangs off (the reference places
        # every node's arrays with numa_alloc_onnode, msb_64.c:2302): run on that node's cores
        # while the page-locked arrays are allocated, so that first touch puts them there
        near = near_cpus(local_rank) if world > 1 else None
        was = os.sched_getaffinity(0) if near else None
        if near:
            try:
                os.sched_setaffinity(0, near)
            except OSError:
                near = None
        hk, hr = m.pinned(cap), m.pinned(cap)
        line_numa = sorted(near)[:1] + sorted(near)[-1:] if near else None
        e2e_s = []
        got = n
        for i in range(args.e2e_steps + 1):
            rc = lib.msb64_b200_memcpy_d2h(hk.ctypes.data, src_k.data_ptr(), n * 8, stream)
            rc |= lib.msb64_b200_memcpy_d2h(hr.ctypes.data, src_r.data_ptr(), n * 8, stream)
            assert rc == 0
            barrier()
            t0 = time.perf_counter()
            if sorter is None:
                size = [n]
                m.sort([hk], [hr], size)             # H2D + sort + D2H, synchronous
            else:
                # host arrays -> this rank's GPU -> sharded sort -> host arrays (this rank's key range)
                rc = lib.msb64_b200_memcpy_h2d(keys.data_ptr(), hk.ctypes.data, n * 8, stream)
                rc |= lib.msb64_b200_memcpy_h2d(rids.data_ptr(), hr.ctypes.data, n * 8, stream)
                assert rc == 0
                ok_, or_, got = sorter.sort(keys, rids)
                rc = lib.msb64_b200_memcpy_d2h(hk.ctypes.data, ok_.data_ptr(), got * 8, stream)
                rc |= lib.msb64_b200_memcpy_d2h(hr.ctypes.data, or_.data_ptr(), got * 8, stream)
                assert rc == 0
                torch.cuda.synchronize()
                dist.barrier()                       # the job is done when every rank has its range back
            dt = time.perf_counter() - t0
            if i:                                # first call is warm-up (allocations)
                e2e_s.append(dt)
        assert bool(np.all(hk[:got - 1] <= hk[1:got])), "e2e output not sorted"
        if sorter is not None:
            # the host arrays hold what the device check just accepted?  compare checksums
            hsum = int(np.sum(hk[:got], dtype=np.uint64))
            _, dsum, _ = check(ok_, or_, got)
            assert hsum == dsum, "e2e host output differs from the device result"
            assert wrap_sum(hsum) == wrap_sum(sum0), "e2e: key checksum over all ranks changed"
        t = torch.tensor([sum(e2e_s) / len(e2e_s)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if was:
            os.sched_setaffinity(0, was)
        line["e2e"] = {"value": world * n / float(t.item()) / 1e9, "unit": UNIT,
                       "host_arrays": ("page-locked, allocated on the cores nearest the rank's GPU (NVML cpu affinity), "
                                       f"rank 0: cpus {line_numa[0]}-{line_numa[1]}" if line_numa else "page-locked"),
                       "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": 16 * n * world,
                       "ms_per_step": float(t.item()) * 1e3, "steps": len(e2e_s),
                       "api": "sort() of include/msb64_b200.h with pinned host arrays" if sorter is None else
                              f"pinned host arrays -> ShardedSorter.sort (exchange = {sorter.exchange}, the sharded "
                              "sort across all ranks) -> pinned host arrays, every rank its key range"}
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
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if sorter is not None:
        sorter.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    if a.ref_child:
        sys.exit(ref_child(parse_count(a.cpu_sample), a.steps))
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
