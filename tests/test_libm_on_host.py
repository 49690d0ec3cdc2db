"""lp_libm.cuh compiled for the host (tests/csrc/libm_harness.cpp): the restated exp / log and the check-free
division of the fit kernel's inner loop, checked against the C library without a GPU.  The hardware
reciprocal seed is emulated here; bit-identity with the CUDA library on the device is asserted by
tests/test_gpu_parity.py::test_device_exp_log_match_the_cuda_library / ::test_device_division_matches_the_cuda_library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "libm_harness.cpp")
HDR = os.path.join(ROOT, "lineagepulse_b200", "csrc", "lp_libm.cuh")
LIB = os.path.join(ROOT, "tests", "_build", "liblibmharness.so")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lm():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", SRC, "-o", LIB])
    lib = C.CDLL(LIB)
    lib.lm_exp.argtypes = [dp, C.c_int, dp, dp]
    lib.lm_log.argtypes = [dp, C.c_int, dp, dp]
    lib.lm_div.argtypes = [dp, dp, C.c_int, dp, dp]
    return lib


def _call2(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    a, b = np.zeros(len(x)), np.zeros(len(x))
    fn(x.ctypes.data_as(dp), len(x), a.ctypes.data_as(dp), b.ctypes.data_as(dp))
    return a, b


def _ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_exp_restatement(lm):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(0, 1, 100000), rng.uniform(-700, 700, 100000), rng.uniform(-745, -700, 20000),
                        rng.uniform(700, 709.7, 20000), [0.0, -0.0, 1.0, -1.0, 709.782712893384, -708.3964185322641]])
    single, paired = _call2(lm.lm_exp, x)
    want = np.exp(x)
    normal = want > 2.3e-308
    assert _ulps(single[normal], want[normal]).max() <= 1.0            # the CUDA library documents 1 ulp for exp
    assert np.all(np.abs(single[~normal] - want[~normal]) <= 5e-324 * 2)   # gradual underflow: within one subnormal step
    assert np.array_equal(single.view(np.uint64), paired.view(np.uint64))  # the pair routine is the same arithmetic
    sp, _ = _call2(lm.lm_exp, np.array([np.inf, -np.inf, np.nan, 710.0, 1000.0, -746.0, -1000.0, 1e308, -1e308]))
    assert sp[0] == np.inf and sp[1] == 0.0 and np.isnan(sp[2])
    assert np.all(sp[3:5] == np.inf) and np.all(sp[5:7] == 0.0) and sp[7] == np.inf and sp[8] == 0.0


def test_log_restatement(lm):
    rng = np.random.default_rng(4)
    x = np.concatenate([np.exp(rng.uniform(-708, 709, 200000)), rng.uniform(0.5, 2.0, 100000),
                        [1.0, np.sqrt(2.0), np.sqrt(0.5), 1.0000000000000002, 0.9999999999999999,
                         2.2250738585072014e-308, 1.7976931348623157e308]])
    full, pos = _call2(lm.lm_log, x)
    want = np.log(x)
    err = np.abs(full - want)
    assert np.all(err <= np.maximum(np.spacing(np.abs(want)), 5e-324))  # 1 ulp
    assert full[len(x) - 7] == 0.0                                     # log(1) is exactly 0
    assert np.array_equal(full.view(np.uint64), pos.view(np.uint64))   # without the special-case tests: same bits
    sub = np.array([5e-324, 1e-310, 2.2250738585072009e-308])
    fs, _ = _call2(lm.lm_log, sub)
    assert _ulps(fs, np.log(sub)).max() <= 1.0                         # subnormal arguments are pre-scaled by 2^54
    sp, _ = _call2(lm.lm_log, np.array([0.0, -0.0, -1.0, np.inf, np.nan, -np.inf]))
    assert sp[0] == -np.inf and sp[1] == -np.inf and np.isnan(sp[2]) and sp[3] == np.inf and np.isnan(sp[4]) and np.isnan(sp[5])


def test_checkfree_division_is_correctly_rounded(lm):
    """Seed + two Newton steps + residual correction: with the seed emulated, the quotient and the reciprocal are
    the correctly rounded IEEE results everywhere in the range the likelihood code uses them."""
    rng = np.random.default_rng(5)
    n = 300000
    for lo, hi in [(-140, 140), (-10, 10), (-0.01, 0.01)]:
        a = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(lo, hi, n) * rng.uniform(1, 2, n)
        b = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(lo, hi, n) * rng.uniform(1, 2, n)
        a[:100] = 0.0
        a[100:200] = b[100:200]
        q, r = np.zeros(n), np.zeros(n)
        lm.lm_div(a.ctypes.data_as(dp), b.ctypes.data_as(dp), n, q.ctypes.data_as(dp), r.ctypes.data_as(dp))
        assert np.array_equal(q, a / b)
        assert np.array_equal(r, 1.0 / b)
    e = np.concatenate([10.0 ** rng.uniform(-320, 300, n), rng.uniform(0, 4, n)])   # the sigmoid's 1 / (1 + e)
    b = 1.0 + e
    q, r = np.zeros(len(b)), np.zeros(len(b))
    lm.lm_div(np.ones_like(b).ctypes.data_as(dp), b.ctypes.data_as(dp), len(b), q.ctypes.data_as(dp), r.ctypes.data_as(dp))
    assert np.array_equal(r, 1.0 / b) and np.array_equal(q, 1.0 / b)
