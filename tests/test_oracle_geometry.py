"""Pins the oracle's collision / offroad / wrong-way geometry against independent implementations:
cv2.rotatedRectangleIntersection, polygon clipping areas and float64 brute-force distances."""
import math

import cv2
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from torchdriveenv_b200 import scenarios as S

EPS_BAND = 1e-4  # metres: cases whose SAT slack is inside the band are excluded (and counted)


def _corners64(b):
    x, y, psi, l, w = [float(v) for v in b]
    c, s = math.cos(psi), math.sin(psi)
    pts = [(l / 2, w / 2), (l / 2, -w / 2), (-l / 2, -w / 2), (-l / 2, w / 2)]
    return np.array([[x + px * c - py * s, y + px * s + py * c] for px, py in pts])


def _clip_area(P, Q):
    """area of the intersection of two convex polygons (Sutherland-Hodgman), float64"""
    def inside(p, a, b): return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0
    def inter(p1, p2, a, b):
        d1, d2 = p2 - p1, b - a
        t = ((a[0] - p1[0]) * d2[1] - (a[1] - p1[1]) * d2[0]) / (d1[0] * d2[1] - d1[1] * d2[0])
        return p1 + t * d1
    def ccw(poly):
        area = 0.5 * sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1] for i in range(len(poly)))
        return poly if area > 0 else poly[::-1]
    out = list(ccw(P)); Q = ccw(Q)
    for i in range(len(Q)):
        a, b = Q[i], Q[(i + 1) % len(Q)]
        inp, out = out, []
        for j in range(len(inp)):
            cur, prv = inp[j], inp[j - 1]
            if inside(cur, a, b):
                if not inside(prv, a, b): out.append(inter(prv, cur, a, b))
                out.append(cur)
            elif inside(prv, a, b):
                out.append(inter(prv, cur, a, b))
        if not out:
            return 0.0
    return abs(0.5 * sum(out[i][0] * out[(i + 1) % len(out)][1] - out[(i + 1) % len(out)][0] * out[i][1] for i in range(len(out))))


def _random_boxes(rng, n, spread):
    return np.stack([rng.uniform(-spread, spread, n), rng.uniform(-spread, spread, n), rng.uniform(-math.pi, math.pi, n),
                     rng.uniform(3.5, 6.0, n), rng.uniform(1.6, 2.4, n)], 1).astype(np.float32)


def test_overlap_vs_clipping_area_and_cv2(oracle):
    rng = np.random.default_rng(0)
    n = 3000
    A, B = _random_boxes(rng, n, 4.0), _random_boxes(rng, n, 4.0)
    st_ = np.zeros((n, 2, 4), np.float32); at = np.ones((n, 2, 4), np.float32)
    st_[:, 0, :3], st_[:, 1, :3] = A[:, :3], B[:, :3]
    at[:, 0, :2], at[:, 1, :2] = A[:, 3:], B[:, 3:]
    margins = oracle.collision_margins(st_, at)[:, 0]
    excluded = 0
    hits = 0
    for k in range(n):
        got = oracle.overlap(A[k], B[k])
        if margins[k] < EPS_BAND:
            excluded += 1
            continue
        area = _clip_area(_corners64(A[k]), _corners64(B[k]))
        assert got == (area > 1e-9), (k, got, area)
        rA = ((float(A[k, 0]), float(A[k, 1])), (float(A[k, 3]), float(A[k, 4])), math.degrees(float(A[k, 2])))
        rB = ((float(B[k, 0]), float(B[k, 1])), (float(B[k, 3]), float(B[k, 4])), math.degrees(float(B[k, 2])))
        ret, _ = cv2.rotatedRectangleIntersection(rA, rB)
        if margins[k] > 1e-2:  # cv2 reports touching as partial, compare away from contact only
            assert got == (ret != cv2.INTERSECT_NONE), (k, got, ret)
        hits += got
    assert excluded < n * 0.01
    assert 0.2 * n < hits < 0.9 * n  # the sample exercises both outcomes


@given(st.floats(-5, 5), st.floats(-5, 5), st.floats(-3.14, 3.14), st.floats(-3.14, 3.14), st.floats(-50, 50), st.floats(-50, 50), st.floats(-3.14, 3.14))
@settings(max_examples=300, deadline=None)
def test_overlap_symmetric_and_rigid_motion_invariant(oracle, dx, dy, pa, pb, tx, ty, rot):
    a = np.array([0, 0, pa, 4.9, 2.0], np.float32)
    b = np.array([dx, dy, pb, 5.2, 2.1], np.float32)
    got = oracle.overlap(a, b)
    assert got == oracle.overlap(b, a)  # bitwise symmetric by construction
    st_ = np.zeros((1, 2, 4), np.float32); at = np.ones((1, 2, 4), np.float32)
    st_[0, 0, :3], st_[0, 1, :3] = a[:3], b[:3]; at[0, 0, :2], at[0, 1, :2] = a[3:], b[3:]
    if oracle.collision_margins(st_, at)[0, 0] < 1e-3:
        return  # inside the band a rigid motion may flip the flag
    c, s = math.cos(rot), math.sin(rot)
    def move(q):
        return np.array([tx + q[0] * c - q[1] * s, ty + q[0] * s + q[1] * c, q[2] + rot, q[3], q[4]], np.float32)
    assert got == oracle.overlap(move(a), move(b))


def test_touching_is_not_collision(oracle):
    a = np.array([0, 0, 0, 4, 2], np.float32)
    assert not oracle.overlap(a, np.array([4, 0, 0, 4, 2], np.float32))      # share an edge
    assert oracle.overlap(a, np.array([3.999, 0, 0, 4, 2], np.float32))
    assert not oracle.overlap(a, np.array([0, 2, 0, 4, 2], np.float32))
    assert oracle.overlap(a, a)


def test_collision_counts(oracle):
    # three boxes in a row, middle one overlaps both neighbours; an absent box is ignored
    st_ = np.zeros((1, 4, 4), np.float32); at = np.ones((1, 4, 4), np.float32)
    at[..., 0], at[..., 1] = 4.0, 2.0
    st_[0, :, 0] = [0, 3, 6, 3]
    at[0, 3, 3] = 0.0
    np.testing.assert_array_equal(oracle.collision_boxes(st_, at)[0], [1, 2, 1, 0])


def _dist64(tris, p):
    best = np.inf
    for t in np.asarray(tris, np.float64):
        v = t[:6].reshape(3, 2)
        cr = [(v[(k + 1) % 3][0] - v[k][0]) * (p[1] - v[k][1]) - (v[(k + 1) % 3][1] - v[k][1]) * (p[0] - v[k][0]) for k in range(3)]
        if all(c >= 0 for c in cr) or all(c <= 0 for c in cr):
            return 0.0
        for k in range(3):
            a, b = v[k], v[(k + 1) % 3]
            ab = b - a
            t_ = np.clip(np.dot(p - a, ab) / np.dot(ab, ab), 0, 1)
            best = min(best, float(np.linalg.norm(p - (a + t_ * ab))))
    return best


def test_point_mesh_distance_vs_float64(oracle):
    m = S.build_polyline_map(S.VALIDATION_POLYLINES["three_way"], "tw")
    rng = np.random.default_rng(3)
    lo, hi = m.road_tris[:, :6].reshape(-1, 2).min(0), m.road_tris[:, :6].reshape(-1, 2).max(0)
    n_in = 0
    for _ in range(400):
        p = rng.uniform(lo - 15, hi + 15)
        p32 = p.astype(np.float32)
        got = oracle.point_mesh_distance(m.road_tris, p32[0], p32[1])
        want = _dist64(m.road_tris, p32.astype(np.float64))
        assert abs(got - want) < 1e-4 * max(1.0, want)
        n_in += want == 0
    assert 20 < n_in < 380
    # a vertex and an edge midpoint are on the mesh
    t = m.road_tris[5]
    assert oracle.point_mesh_distance(m.road_tris, t[0], t[1]) == 0.0
    assert oracle.point_mesh_distance(m.road_tris, 0.5 * (t[0] + t[2]), 0.5 * (t[1] + t[3])) < 1e-5


def test_offroad_threshold_semantics(oracle):
    # one square road tile [0,10]^2 made of two triangles
    tris = np.array([[0, 0, 10, 0, 10, 10, 1, 0], [0, 0, 10, 10, 0, 10, 1, 0]], np.float32)
    def off(x, y, psi=0.0, l=4.0, w=2.0, thr=0.5):
        st_ = np.array([[[x, y, psi, 0]]], np.float32); at = np.array([[[l, w, 1, 1]]], np.float32)
        return float(oracle.offroad_boxes(tris, thr, st_, at)[0, 0])
    assert off(5, 5) == 0.0                       # fully on the road
    assert off(8.4, 5) == 0.0                     # corners at x = 10.4: 0.4 m off < 0.5 m threshold
    v = off(9.0, 5)                               # corners at x = 11: two corners 1.0 m off -> 2 * 0.5
    assert abs(v - 1.0) < 1e-5
    assert abs(off(9.0, 5, thr=0.0) - 2.0) < 1e-5  # threshold 0 degenerates to plain distance (point-in-mesh test)
    far = off(30, 5)                              # every corner off: distances 18,18,22,22 minus 0.5 each
    assert abs(far - (17.5 * 2 + 21.5 * 2)) < 1e-3
