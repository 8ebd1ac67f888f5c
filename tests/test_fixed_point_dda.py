"""The arithmetic of the rasteriser's edge accumulators (edge_fixed + the row loop of raster_batch in
torchdriveenv_b200/csrc/tde_render.cuh), restated operation by operation against exact integer floors.

Claim under test: with r ~ 2^32 / d from a binary32 reciprocal seed refined by two FP64 Newton steps,
acc = floor(K r) + 2^13 and inc = floor(S r), the high word of acc + i * inc equals floor((K + S i) / d) for every row
i <= 63, every d = 16 |dy| < 2^18 and every K, S a snapped primitive can produce - also when the seed is off by a few
binary32 ulps (MUFU.RCP is not correctly rounded).  The fused multiply-adds are evaluated exactly (rational arithmetic,
one rounding), as the hardware does.  The GPU tests check pixels; this checks the number theory on all of the domain."""
from fractions import Fraction

import numpy as np
import pytest


def fma(a: float, b: float, c: float) -> float:
    return float(Fraction(a) * Fraction(b) + Fraction(c))   # float(Fraction) rounds to nearest even


def edge_fixed(K: int, S: int, d: int, seed_ulps: int):
    """edge_fixed(): returns (acc, inc) as Python ints (two's complement not applied: plain integers)."""
    rf = np.float32(1.0) / np.float32(d)
    for _ in range(abs(seed_ulps)):
        rf = np.nextafter(rf, np.float32(np.inf) if seed_ulps > 0 else np.float32(-np.inf))
    dd, r = float(d), float(rf)
    e = fma(-dd, r, 1.0)
    r = fma(r, e, r)
    e = fma(-dd, r, 1.0)
    r = fma(r, e, r)
    r = r * 4294967296.0
    acc = int(np.floor(float(K) * r)) + 8192      # one rounding in the product, then floor (F2I.S64.F64.FLOOR)
    inc = int(np.floor(float(S) * r))
    return acc, inc


def check(K, S, d, seed_ulps):
    for k, s, dd in zip(K, S, d):
        k, s, dd = int(k), int(s), int(dd)
        acc, inc = edge_fixed(k, s, dd, seed_ulps)
        for i in (0, 1, 2, 3, 7, 15, 31, 32, 47, 62, 63):
            got = (acc + i * inc) >> 32
            want = (k + s * i) // dd
            assert got == want, f"row {i}: K={k} S={s} d={dd} seed_ulps={seed_ulps}: got {got} want {want}"


@pytest.mark.parametrize("ulps", [0, 2, -2, 4, -4])
def test_edge_accumulators_are_exact_on_random_edges(ulps):
    rng = np.random.default_rng(10 + ulps)
    n = 6000
    dy = rng.integers(1, 16383, n); dy[: n // 4] = rng.integers(1, 40, n // 4); dy[n // 4: n // 2] = rng.integers(16000, 16383, n // 4)
    d = 16 * dy
    S = 16 * rng.integers(-16382, 16383, n)
    K = rng.integers(-600_000_000, 600_000_001, n)     # |K| < 2^30 covers every snapped primitive (|C0| < 2^29)
    check(K, S, d, ulps)


def test_edge_accumulators_are_exact_on_boundary_cases():
    """Remainders 0 and d - 1, slopes whose fraction is 0, 1/d and (d-1)/d, the largest d, exact multiples."""
    ds, Ks, Ss = [], [], []
    for d in (16, 32, 48, 16 * 255, 16 * 256, 16 * 257, 16 * 8191, 16 * 8192, 16 * 16381, 16 * 16382):
        for k_rem in (0, 1, d // 2, d - 2, d - 1):
            for k_q in (-70_000, -1, 0, 1, 1023, 70_000):
                for s in (0, 16, -16, d, -d, d - 16, -(d - 16), d + 16, 16 * 16382, -16 * 16382, (d // 32) * 16):
                    ds.append(d); Ks.append(k_q * d + k_rem); Ss.append(s)
    # the largest quotients a drawn primitive can reach: |K| ~ 2^29 over the smallest d
    for k in (536_870_000, -536_870_000, 536_870_911, -536_870_912):
        for d in (16, 32, 16 * 16382):
            ds.append(d); Ks.append(k); Ss.append(16 * 16382)
    for ulps in (0, 3, -3):
        check(Ks, Ss, ds, ulps)
