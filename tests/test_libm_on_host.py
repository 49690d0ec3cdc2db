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


def test_integer_range_tests_select_like_the_fp64_compares():
    """The fit kernel decides three range tests on the upper 32-bit word instead of with fp64 compares (lp_math.cuh
    lp_div_safe and lp_bd0_t, lp_fit_kernel.cuh FastLane::mean; profiles/README.md R2.12).  Each only selects between two
    paths that return the same bits where both apply, so what has to hold is the implication integer test => fp64
    condition (and equivalence for bd0's early exit).  Constants are read off the sources so the test follows them."""
    src = {n: open(os.path.join(ROOT, "lineagepulse_b200", "csrc", n)).read() for n in ("lp_math.cuh", "lp_fit_kernel.cuh")}
    assert "lo = 0x2b300000u, span = 0x54a00000u - 0x2b300000u" in src["lp_math.cuh"]
    assert "(__double2hiint(s) & 0x7ff00000) == 0" in src["lp_math.cuh"]
    assert "(unsigned)__double2hiint(e) < 0x7e37e43cu" in src["lp_fit_kernel.cuh"]

    def hi(x):
        return (np.asarray(x, dtype=np.float64).view(np.uint64) >> np.uint64(32)).astype(np.uint32)

    def from_hi_lo(h, lo):
        return ((np.asarray(h, dtype=np.uint64) << np.uint64(32)) | np.asarray(lo, dtype=np.uint64)).view(np.float64)

    rng = np.random.default_rng(11)
    # all exponents and signs, random mantissas, plus the words around every constant with extreme lower words
    words = np.concatenate([rng.integers(0, 1 << 32, 400000, dtype=np.uint64),
                            np.array([0x2b2fffff, 0x2b300000, 0x2b300001, 0x549fffff, 0x54a00000, 0x54a00001, 0x7e37e43b,
                                      0x7e37e43c, 0x7e37e43d, 0x000fffff, 0x00100000, 0x800fffff, 0x80100000, 0, 0x80000000,
                                      0x7ff00000, 0x7ff80000, 0xfff00000, 0xfff80000], dtype=np.uint64)])
    for lo_word in (0, 0xffffffff, 1):
        v = from_hi_lo(words, np.full(len(words), lo_word, dtype=np.uint64))
        h = hi(v)
        with np.errstate(invalid="ignore"):
            # lp_div_safe: [2^-332, 2^331) lies inside (1e-100, 1e100); negative, NaN, infinite operands fail
            int_safe = (h - np.uint32(0x2b300000)) < np.uint32(0x54a00000 - 0x2b300000)
            fp_safe = (v > 1e-100) & (v < 1e100)
            assert np.all(fp_safe[int_safe])
            assert not np.any(int_safe & ~np.isfinite(v)) and not np.any(int_safe & (v <= 0))
            # the sigmoid's reciprocal: e >= 0 or NaN arrives; the integer test must imply e < 1e300
            int_small = h < np.uint32(0x7e37e43c)
            assert np.all((v < 1e300)[int_small & (v >= 0)])
            assert not np.any(int_small & np.isnan(v))
            assert not np.any(int_small & (v < 0) & (v != 0))   # -0.0 cannot arrive from exp(); negatives never pass
            # bd0's subnormal exit: exponent field zero <=> |s| < 2^-1022 (DBL_MIN), zero included, NaN excluded
            int_sub = (h & np.uint32(0x7ff00000)) == 0
            assert np.array_equal(int_sub, np.abs(v) < 2.2250738585072014e-308)
    assert 2.0 ** -332 > 1e-100 and 2.0 ** 331 < 1e100 and hi(1e300) == 0x7e37e43c
