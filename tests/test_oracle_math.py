"""Pins the oracle's arithmetic against independent float64 closed forms (the reference ships no
golden vectors)."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def _ulp_err(a32, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(a32.astype(np.float64) - ref64) / np.maximum(ulp, 1e-45)


def test_sincos_accuracy(oracle):
    rng = np.random.default_rng(0)
    for span in (0.5, 4.0, 100.0, 1e4):
        x = rng.uniform(-span, span, 200_000).astype(np.float32)
        s, c = oracle.sincos(x)
        x64 = x.astype(np.float64)
        # absolute error bound (values near zero crossings have tiny ulps)
        assert np.abs(s - np.sin(x64)).max() < 2.5e-7 * max(1.0, span / 100)
        assert np.abs(c - np.cos(x64)).max() < 2.5e-7 * max(1.0, span / 100)
    x = rng.uniform(-3.5, 3.5, 500_000).astype(np.float32)
    s, c = oracle.sincos(x)
    big = np.abs(np.sin(x.astype(np.float64))) > 1e-2
    assert _ulp_err(s, np.sin(x.astype(np.float64)))[big].max() < 2.0
    big = np.abs(np.cos(x.astype(np.float64))) > 1e-2
    assert _ulp_err(c, np.cos(x.astype(np.float64)))[big].max() < 2.0


def test_sincos_special_values(oracle):
    s, c = oracle.sincos(np.array([0.0, -0.0], np.float32))
    assert s[0] == 0.0 and c[0] == 1.0 and c[1] == 1.0
    s, c = oracle.sincos(np.array([np.pi / 2, np.pi, -np.pi / 2], np.float32))
    assert abs(s[0] - 1) < 1e-7 and abs(c[1] + 1) < 1e-7 and abs(s[2] + 1) < 1e-7
    x = np.linspace(-10, 10, 4001).astype(np.float32)
    s, c = oracle.sincos(x)
    assert np.abs(s * s + c * c - 1).max() < 4e-7


@given(st.floats(min_value=-50, max_value=50, allow_nan=False, width=32))
@settings(max_examples=300, deadline=None)
def test_wrap_pi_matches_python_floored_modulo(oracle, psi):
    # ((pi + psi) mod 2pi) - pi with Python's floored modulo, evaluated in binary32 like torch does
    f = np.float32
    t = f(f(psi) + f(np.pi))
    m = f(np.fmod(t, f(2 * np.pi)))
    if m < 0:
        m = f(m + f(2 * np.pi))
    want = f(m - f(np.pi))
    got = oracle.wrap_pi(psi)
    assert got == float(want)
    assert -math.pi - 1e-6 <= got < math.pi + 1e-6
    # and agrees with the real-number definition
    ref = (math.pi + psi) % (2 * math.pi) - math.pi
    d = abs(got - ref)
    assert min(d, abs(d - 2 * math.pi)) < 1e-5


def _bicycle64(state, action, lr, dt=0.1):
    """float64 restatement of SURVEY.md §8a a2 (KinematicBicycle.step)."""
    x, y, psi, v = [np.asarray(state[..., k], np.float64) for k in range(4)]
    a, beta = np.asarray(action[..., 0], np.float64), np.asarray(action[..., 1], np.float64)
    v = v + a * dt
    x = x + v * np.cos(psi + beta) * dt
    y = y + v * np.sin(psi + beta) * dt
    psi = psi + (v / lr) * np.sin(beta) * dt
    psi = (np.pi + psi) % (2 * np.pi) - np.pi
    return np.stack([x, y, psi, v], -1)


def test_bicycle_vs_float64(oracle):
    rng = np.random.default_rng(1)
    n = 20000
    state = np.stack([rng.uniform(-200, 200, n), rng.uniform(-200, 200, n), rng.uniform(-3.1, 3.1, n), rng.uniform(-5, 15, n)], 1).astype(np.float32)
    action = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.3, 0.3, n)], 1).astype(np.float32)
    lr = rng.uniform(0.8, 2.5, n).astype(np.float32)
    got = oracle.bicycle_step(state, action, lr)
    want = _bicycle64(state, action, lr.astype(np.float64))
    assert np.abs(got[:, :2] - want[:, :2]).max() < 1e-5 * 200
    dpsi = np.abs(got[:, 2] - want[:, 2]); dpsi = np.minimum(dpsi, np.abs(dpsi - 2 * np.pi))
    assert dpsi.max() < 2e-6
    assert np.abs(got[:, 3] - want[:, 3]).max() < 2e-6
    assert (got[:, 2] >= -np.pi - 1e-6).all() and (got[:, 2] < np.pi + 1e-6).all()


def test_bicycle_straight_line_constant_velocity(oracle):
    # beta = 0, a = 0: straight line at constant speed, heading unchanged (up to the wrap's rounding)
    state = np.array([[10.0, -5.0, 0.7, 8.0]], np.float32)
    s = state.copy()
    for _ in range(200):
        s = oracle.bicycle_step(s, np.zeros((1, 2), np.float32), np.array([1.0], np.float32))
    assert abs(s[0, 3] - 8.0) == 0.0
    assert abs(s[0, 2] - 0.7) < 1e-5
    assert abs(s[0, 0] - (10 + 160 * math.cos(0.7))) < 2e-3 and abs(s[0, 1] - (-5 + 160 * math.sin(0.7))) < 2e-3


def test_bicycle_trajectory_200_steps(oracle):
    rng = np.random.default_rng(2)
    n = 256
    s32 = np.stack([rng.uniform(-100, 100, n), rng.uniform(-100, 100, n), rng.uniform(-3, 3, n), rng.uniform(0, 10, n)], 1).astype(np.float32)
    s64 = s32.astype(np.float64)
    lr = rng.uniform(0.82, 0.97, n).astype(np.float32)
    for _ in range(200):
        a = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.3, 0.3, n)], 1).astype(np.float32)
        s32 = oracle.bicycle_step(s32, a, lr)
        s64 = _bicycle64(s64, a.astype(np.float64), lr.astype(np.float64))
    scale = np.maximum(1.0, np.abs(s64[:, :2]).max())
    assert np.abs(s32[:, :2] - s64[:, :2]).max() / scale < 1e-4   # 200 accumulated binary32 steps
    assert np.abs(s32[:, 3] - s64[:, 3]).max() < 1e-4


def test_rng_is_counter_based(oracle):
    L = oracle.lib()
    a = L.orc_rng(1, 2, 3, 4)
    assert a == L.orc_rng(1, 2, 3, 4)
    assert len({L.orc_rng(1, 2, 3, k) for k in range(64)}) == 64
    assert len({L.orc_rng(s, 0, 0, 0) for s in range(64)}) == 64
