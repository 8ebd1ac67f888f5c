"""The arithmetic of the rasteriser's edge accumulators (cover_rows64 / tde_frac32 / tde_floordiv in
torchdriveenv_b200/csrc), restated operation by operation in numpy (binary32 where the kernel uses binary32), against
exact integer floors.  Claim under test: with acc = floor(K/d) * 2^32 + frac32(K mod d) + 2^13 and
inc = floor(S/d) * 2^32 + frac32(S mod d), the high word of acc + i * inc equals floor((K + S i) / d) for every row
i <= 63, every d = 16 |dy| < 2^18 and every K, S a snapped primitive can produce - also when the reciprocal estimate
of d is off by a few ulps (__fdividef).  The GPU tests check pixels; this checks the number theory on all of the domain."""
import numpy as np
import pytest

F = np.float32
U32 = np.uint64(0xFFFFFFFF)


def floordiv(a, d, inv_d):
    """tde_floordiv: float estimate refined twice in integers, then one fix-up; a, d int64 arrays, inv_d float32."""
    q = np.floor(a.astype(F) * inv_d).astype(np.int64)
    r = a - q * d
    q2 = np.floor(r.astype(F) * inv_d).astype(np.int64)
    q = q + q2
    r = r - q2 * d
    lo, hi = r < 0, r >= d
    q = q - lo + hi
    r = r + lo * d - hi * d
    return q, r


def frac32(r, d, inv_d):
    """tde_frac32: floor(r * 2^32 / d) to within +-1 for 0 <= r < d."""
    b1 = np.trunc(r.astype(F) * (inv_d * F(4294967296.0))).astype(np.uint64)
    assert (b1 <= U32).all()
    low = (b1 * d.astype(np.uint64)) & U32                       # b1 * (unsigned)d, 32-bit wrap
    res = -(low.astype(np.int64) - ((low >> np.uint64(31)) << np.uint64(32)).astype(np.int64))   # -(int)low
    corr = np.rint(res.astype(F) * inv_d).astype(np.int64)
    return ((b1.astype(np.int64) + corr) & 0xFFFFFFFF).astype(np.uint64)


def check(K, S, d, ulps):
    K, S, d = (np.asarray(v, np.int64) for v in (K, S, d))
    inv_d = (F(1.0) / d.astype(F))
    inv_d = np.nextafter(inv_d, F(np.inf) if ulps > 0 else F(-np.inf)) if ulps else inv_d
    for _ in range(max(0, abs(ulps) - 1)):
        inv_d = np.nextafter(inv_d, F(np.inf) if ulps > 0 else F(-np.inf))
    f0, r0 = floordiv(K, d, inv_d)
    qs, rs = floordiv(S, d, inv_d)
    assert (f0 == K // d).all() and (r0 == K % d).all() and (qs == S // d).all() and (rs == S % d).all(), "tde_floordiv is not exact"
    a_frac, b_frac = frac32(r0, d, inv_d), frac32(rs, d, inv_d)
    # exact values for comparison: |frac32 - r * 2^32 / d| <= 1 (+ rounding)
    for got, r in ((a_frac, r0), (b_frac, rs)):
        exact = (r.astype(object) << 32) // d.astype(object)
        err = np.abs(got.astype(object) - exact)
        assert max(err) <= 2, max(err)
    assert (a_frac + np.uint64(8192) <= U32).all(), "the biased fraction must not carry into the integer part"
    acc = (f0.astype(object) << 32) + (a_frac.astype(object) + 8192)
    inc = (qs.astype(object) << 32) + b_frac.astype(object)
    Ko, So, do = K.astype(object), S.astype(object), d.astype(object)
    for i in (0, 1, 2, 3, 7, 15, 31, 32, 47, 62, 63):
        want = (Ko + So * i) // do
        got = (acc + inc * i) >> 32
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, f"row {i}: K={K[bad[0]]} S={S[bad[0]]} d={d[bad[0]]} got {got[bad[0]]} want {want[bad[0]]}"


@pytest.mark.parametrize("ulps", [0, 2, -2, 4, -4])
def test_edge_accumulators_are_exact_on_random_edges(ulps):
    rng = np.random.default_rng(10 + ulps)
    n = 120_000
    dy = rng.integers(1, 16383, n); dy[: n // 4] = rng.integers(1, 40, n // 4); dy[n // 4: n // 2] = rng.integers(16000, 16383, n // 4)
    d = 16 * dy
    S = 16 * rng.integers(-16382, 16383, n)
    K = rng.integers(-300_000_000, 300_000_001, n)
    check(K, S, d, ulps)


def test_edge_accumulators_are_exact_on_boundary_cases():
    """Remainders 0 and d - 1, slopes whose fraction is 0, 1/d and (d-1)/d, the largest d, exact multiples."""
    ds, Ks, Ss = [], [], []
    for d in (16, 32, 48, 16 * 3, 16 * 255, 16 * 256, 16 * 257, 16 * 8191, 16 * 8192, 16 * 16381, 16 * 16382):
        for k_rem in (0, 1, d // 2, d - 2, d - 1):
            for k_q in (-70_000, -1, 0, 1, 1023, 70_000):
                for s in (0, 16, -16, d, -d, d - 16, -(d - 16), d + 16, 16 * 16382, -16 * 16382, (d // 32) * 16):
                    ds.append(d); Ks.append(k_q * d + k_rem); Ss.append(s)
    for ulps in (0, 3, -3):
        check(Ks, Ss, ds, ulps)
