"""The oracle's float32 elementary functions (restated chewxy/math32) against float64 libm: they must be accurate to
a couple of float32 ulps, otherwise the restatement (not just its last bit) is wrong."""
import ctypes as C
import math

import numpy as np
import pytest


def ulp_err(got, want64):
    want32 = np.float32(want64)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    ulp = np.maximum(ulp, np.float64(np.finfo(np.float32).tiny))
    return np.abs(got.astype(np.float64) - want64) / ulp


def vec1(fn, xs):
    return np.array([fn(C.c_float(float(x))) for x in xs], dtype=np.float32)


def vec2(fn, xs, ys):
    return np.array([fn(C.c_float(float(x)), C.c_float(float(y))) for x, y in zip(xs, ys)], dtype=np.float32)


@pytest.fixture(scope="module")
def L(oracle):
    return oracle.lib()


def test_sin_cos_tan_accuracy(L):
    """float32 three-part pi/4 reduction: ~2 ulp inside one period, degrading slowly with |x| (y*PI4A is no longer
    exact once y needs more than a couple of bits) -- the same behaviour a float32 port of Go's math.Sin has."""
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-2 * math.pi, 2 * math.pi, 4000), np.linspace(-6.28, 6.28, 1401)]).astype(np.float32)
    assert np.abs(vec1(L.go_sin, xs).astype(np.float64) - np.sin(xs.astype(np.float64))).max() < 3e-7
    assert np.abs(vec1(L.go_cos, xs).astype(np.float64) - np.cos(xs.astype(np.float64))).max() < 3e-7
    xw = rng.uniform(-40, 40, 3000).astype(np.float32)
    assert np.abs(vec1(L.go_sin, xw).astype(np.float64) - np.sin(xw.astype(np.float64))).max() < 3e-6
    assert np.abs(vec1(L.go_cos, xw).astype(np.float64) - np.cos(xw.astype(np.float64))).max() < 3e-6
    xt = rng.uniform(-1.2, 1.2, 3000).astype(np.float32)
    assert ulp_err(vec1(L.go_tan, xt), np.tan(xt.astype(np.float64))).max() <= 3


def test_sin_cos_special(L):
    assert L.go_sin(0.0) == 0.0 and L.go_cos(0.0) == 1.0
    assert math.copysign(1, L.go_sin(-0.0)) == -1
    assert math.isnan(L.go_sin(float("inf"))) and math.isnan(L.go_cos(float("nan")))
    assert L.go_tan(0.0) == 0.0
    # standard NPT taper: tan(atan(1/32)) ~ 1/32
    assert abs(L.go_tan(L.go_atan(1.0 / 32.0)) - 1.0 / 32.0) < 1e-8


def test_atan_atan2_accuracy(L):
    rng = np.random.default_rng(11)
    xs = np.concatenate([rng.uniform(-5, 5, 3000), rng.uniform(-1e3, 1e3, 500), [0.66, 0.6600001, 2.4142135, 2.4142137]]).astype(np.float32)
    assert ulp_err(vec1(L.go_atan, xs), np.arctan(xs.astype(np.float64))).max() <= 2
    ys = rng.uniform(-3, 3, 4000).astype(np.float32)
    xx = rng.uniform(-3, 3, 4000).astype(np.float32)
    got = vec2(L.go_atan2, ys, xx)
    assert np.abs(got.astype(np.float64) - np.arctan2(ys.astype(np.float64), xx.astype(np.float64))).max() < 5e-7


def test_atan2_special(L):
    pi = np.float32(math.pi)
    assert L.go_atan2(0.0, 1.0) == 0.0
    assert np.float32(L.go_atan2(0.0, -1.0)) == pi
    assert np.float32(L.go_atan2(-0.0, -1.0)) == -pi
    assert np.float32(L.go_atan2(1.0, 0.0)) == np.float32(math.pi / 2)
    assert np.float32(L.go_atan2(-1.0, 0.0)) == -np.float32(math.pi / 2)
    assert math.isnan(L.go_atan2(float("nan"), 1.0))


def test_hypot_is_scaled_form(L):
    """math32.Hypot = p*sqrt(1+(q/p)^2): accurate, overflow safe, and hypot(0,0)=0."""
    rng = np.random.default_rng(3)
    a = rng.uniform(-100, 100, 3000).astype(np.float32)
    b = rng.uniform(-100, 100, 3000).astype(np.float32)
    got = vec2(L.go_hypot, a, b)
    assert ulp_err(got, np.hypot(a.astype(np.float64), b.astype(np.float64))).max() <= 2
    assert L.go_hypot(0.0, 0.0) == 0.0
    assert L.go_hypot(3e30, 4e30) == pytest.approx(5e30, rel=1e-6)  # sqrt(p*p+q*q) would overflow float32
    assert L.go_hypot(float("inf"), 1.0) == float("inf")
    # exact restatement: p * sqrt(1 + (q/p)^2) in float32
    p, q = np.float32(3.7), np.float32(1.3)
    r = np.float32(q / p)
    want = np.float32(p * np.sqrt(np.float32(np.float32(1) + np.float32(r * r))))
    assert np.float32(L.go_hypot(1.3, 3.7)) == want


def test_min_max_go_semantics(L):
    assert math.isnan(L.go_min(float("nan"), 1.0)) and math.isnan(L.go_max(1.0, float("nan")))
    assert math.copysign(1, L.go_min(0.0, -0.0)) == -1 and math.copysign(1, L.go_max(-0.0, 0.0)) == 1
    assert L.go_min(float("-inf"), float("nan")) == float("-inf")
    assert L.go_min(2.0, 1.0) == 1.0 and L.go_max(2.0, 1.0) == 2.0


def test_floor_round(L):
    assert L.go_floor(-0.5) == -1.0 and L.go_floor(2.9999) == 2.0
    assert L.go_round(0.5) == 1.0 and L.go_round(-0.5) == -1.0 and L.go_round(2.5) == 3.0  # half away from zero
