"""Seeded input families shared by the parity tests (the cases the reference's own
benchmark and BASELINE.json configs exercise: uniform, few distinct values, low bits
only, presorted, reverse sorted, all equal, adversarial bit patterns)."""
import numpy as np


def make(kind: str, n: int, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    u = lambda: rng.integers(0, 1 << 64, size=n, dtype=np.uint64)  # noqa: E731
    if kind == "uniform":
        return u()
    if kind == "low24":
        return u() & np.uint64((1 << 24) - 1)
    if kind == "low8":
        return u() & np.uint64(0xFF)
    if kind == "high8":
        return u() & np.uint64(0xFF << 56)
    if kind == "dup16":            # 16 distinct values spread over the key space
        vals = rng.integers(0, 1 << 64, size=16, dtype=np.uint64)
        return vals[rng.integers(0, 16, size=n)]
    if kind == "dup1000":
        vals = rng.integers(0, 1 << 64, size=1000, dtype=np.uint64)
        return vals[rng.integers(0, 1000, size=n)]
    if kind == "dupmid":           # ~700 copies of each value: long all-equal bins inside local-sort units
        vals = rng.integers(0, 1 << 64, size=max(n // 700, 1), dtype=np.uint64)
        return vals[rng.integers(0, vals.size, size=n)]
    if kind == "duppair":          # long bins holding two different keys each (not all equal)
        vals = rng.integers(0, 1 << 64, size=max(n // 1400, 1), dtype=np.uint64) & ~np.uint64(1)
        return vals[rng.integers(0, vals.size, size=n)] | rng.integers(0, 2, size=n).astype(np.uint64)
    if kind == "dup36":            # ~36 copies of each value: long bins, some holding two values
        vals = rng.integers(0, 1 << 64, size=max(n // 36, 1), dtype=np.uint64)
        return vals[rng.integers(0, vals.size, size=n)]
    if kind == "equal":
        return np.full(n, 0xDEADBEEFCAFEF00D, dtype=np.uint64)
    if kind == "zero":
        return np.zeros(n, dtype=np.uint64)
    if kind == "ones":
        return np.full(n, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    if kind == "sorted":
        return np.sort(u())
    if kind == "reverse":
        return np.sort(u())[::-1].copy()
    if kind == "iota":
        return np.arange(n, dtype=np.uint64)
    if kind == "iota_rev":
        return np.arange(n, dtype=np.uint64)[::-1].copy()
    if kind == "skew":             # 90% of the keys in one top-digit bucket
        a = u()
        heavy = rng.random(n) < 0.9
        a[heavy] = (a[heavy] >> np.uint64(8)) | np.uint64(0x42 << 56)
        return a
    if kind == "zipf":
        z = rng.zipf(1.3, size=n).astype(np.uint64)
        return z * np.uint64(0x9E3779B97F4A7C15)
    if kind == "pow2":             # one bit set: every digit position is hit by few keys
        return np.uint64(1) << rng.integers(0, 64, size=n).astype(np.uint64)
    if kind == "outlier":          # one huge key, the rest differ in low bits only
        a = u() & np.uint64(0xFFF)
        if n:
            a[rng.integers(0, n)] = np.uint64(1 << 63)
        return a
    if kind == "midbits":          # entropy only in bits 20..39
        return (u() & np.uint64((1 << 20) - 1)) << np.uint64(20)
    if kind == "two":
        return rng.integers(0, 2, size=n).astype(np.uint64) * np.uint64(1 << 40)
    if kind == "clustered":        # 4096-key clusters sharing 52 high bits, differing in low 12
        base = rng.integers(0, 1 << 64, size=max(n // 3000, 1), dtype=np.uint64) & ~np.uint64(0xFFF)
        return base[rng.integers(0, base.size, size=n)] | (u() & np.uint64(0x3))
    raise ValueError(kind)


KINDS = ["uniform", "low24", "low8", "high8", "dup16", "dup1000", "dupmid", "duppair", "dup36", "equal", "zero", "ones",
         "sorted", "reverse", "iota", "iota_rev", "skew", "zipf", "pow2", "outlier", "midbits",
         "two", "clustered"]
