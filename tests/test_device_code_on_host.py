"""CPU-only checks of the code the kernels run: lp_math.cuh / lp_model.cuh / lp_bfgs.cuh are
LP_HD (host + device) and are compiled here with g++ (tests/csrc/host_harness.cpp).

  * the resumable vmmin state machine takes exactly the decisions of the oracle's literal vmmin
    restatement (identical counts on Rosenbrock; bit-identical fits when fed identical objectives);
  * the device per-observation likelihood agrees with mpmath and with the oracle to ~1e-15.
"""
import ctypes as C

import numpy as np
import pytest

from lineagepulse_b200 import _cabi as A
from tests.helpers import design_for, golden, model
from tests.host_harness import load


@pytest.fixture(scope="module")
def hh():
    return load()


def test_state_machine_rosenbrock(hh, oracle):
    par = np.zeros(2)
    val = C.c_double()
    fc, gc, cv = C.c_int(), C.c_int(), C.c_int()
    hh.hh_rosenbrock(100, np.sqrt(np.finfo(float).eps), par.ctypes.data_as(A.c_double_p), C.byref(val),
                     C.byref(fc), C.byref(gc), C.byref(cv))
    ref = oracle.rosenbrock(use_numgrad=True)
    assert (fc.value, gc.value, cv.value) == (ref["fncount"], ref["grcount"], ref["convergence"])
    assert val.value == ref["value"] and np.array_equal(par, ref["par"])


def test_device_obs_math_vs_mpmath_and_oracle(hh, oracle):
    worst = 0.0
    for r in golden("zinb_obs.json"):
        v = hh.hh_ll_obs(r["x"], r["mu"], r["size"], r["pi"])
        tol = 1e-13 if r["size"] <= 1e4 else 1e-5
        assert abs(v - r["value"]) <= tol * max(1.0, abs(r["value"])), r
        worst = max(worst, abs(v - oracle.ll_obs(r["x"], r["mu"], r["size"], r["pi"])) / max(1.0, abs(r["value"])))
    assert worst < 5e-15
    for r in golden("dnbinom_mu_log.json"):
        if r["value"] < -700:
            continue
        v = hh.hh_ll_obs(r["x"], r["mu"], r["size"], -1.0)
        o = oracle.dnbinom_mu_log(r["x"], r["size"], r["mu"])
        # same algorithm as R's dnbinom_mu => the two agree far beyond their distance to mpmath
        assert abs(v - o) <= 1e-13 * max(1.0, abs(o)), r


def test_dnbinom_beyond_the_dispersion_clamp(hh, oracle):
    """phi > 1e10 (dispersion batch factors, spline dispersions): R's dnbinom_mu switches to its
    x < 1e-10 * size formula; the device code and the oracle follow (R/srcLineagePulse_evalLogLik.R:136-140)."""
    n_branch = 0
    for r in golden("dnbinom_huge_size.json"):
        v = hh.hh_ll_obs(r["x"], r["mu"], r["size"], -1.0)
        o = oracle.dnbinom_mu_log(r["x"], r["size"], r["mu"])
        want = max(r["value"], -700.0)
        small_x = r["x"] < 1e-10 * r["size"]
        n_branch += small_x
        # the small-x formula drops O(mu x / size) terms (2e-9 at size 1.5e10); the dbinom_raw branch keeps R's
        # size * 1e-17 digit loss
        tol = 5e-9 if small_x else 1e-5
        assert abs(max(o, -700.0) - want) <= tol * max(1.0, abs(want)), r
        assert abs(v - max(o, -700.0)) <= 1e-13 * max(1.0, abs(o)), r
    assert n_branch >= 60


def _fit_gene(hh, d, m, X, i, th0, drop, pi_fixed, use_oracle_obs):
    hh.hh_use_oracle_obs(int(use_oracle_obs))
    out = np.zeros(len(th0))
    ll, cv, ev = C.c_double(), C.c_int(), C.c_uint()
    hh.hh_fit_gene(C.byref(d.struct), C.byref(m), X[i].ctypes.data_as(A.c_int32_p), i,
                   np.ascontiguousarray(th0).ctypes.data_as(A.c_double_p),
                   drop.ctypes.data_as(A.c_double_p) if drop is not None else None,
                   pi_fixed.ctypes.data_as(A.c_double_p) if pi_fixed is not None else None,
                   1000, 1e-8, 1e-3, out.ctypes.data_as(A.c_double_p), C.byref(ll), C.byref(cv), C.byref(ev))
    hh.hh_use_oracle_obs(0)
    return out, ll.value, cv.value, ev.value


def test_state_machine_equals_oracle_vmmin_on_gene_fits(hh, oracle, c1):
    """With bit-identical objective values the resumable state machine + theta transforms must
    reproduce the oracle's fits bit for bit (impulse ZINB, all three starts, every 9th gene)."""
    X = c1["X"]
    d = design_for(c1)
    m = model("impulse", drop="logistic")
    drop = np.full((c1["N"], 1), -2.2, order="F")
    p0 = oracle.init_params(X, d, m)
    p0.mat_drop = drop.copy()
    pk, vl = oracle.impulse_starts(X, d)
    pO = p0.copy()
    rO = oracle.fit_mudisp(X, d, m, pO)
    pi_fixed = 0.001 + 0.998 / (1 + np.exp(-drop[:, 0]))
    for i in range(0, c1["G"], 9):
        prior = p0.mat_mu[i].copy()
        prior[2:5] = np.log(prior[2:5])
        lls = []
        for start in (pk[i], vl[i], prior):
            th0 = np.r_[np.log(p0.mat_disp[i, 0]), start]
            _, ll, cv, _ = _fit_gene(hh, d, m, X, i, th0, drop, pi_fixed, use_oracle_obs=True)
            lls.append(ll)
        assert max(lls) == rO["ll"][i]


def test_device_objective_constant_model_fits_match_oracle(hh, oracle, c1):
    """Well-conditioned fits (constant mean, NB) agree tightly even with the device arithmetic."""
    X = c1["X"]
    d = design_for(c1)
    m = model("constant")
    p0 = oracle.init_params(X, d, m)
    pO = p0.copy()
    rO = oracle.fit_mudisp(X, d, m, pO)
    for i in range(0, c1["G"], 5):
        th0 = np.r_[np.log(p0.mat_disp[i, 0]), np.log(p0.mat_mu[i, 0])]
        th, ll, cv, _ = _fit_gene(hh, d, m, X, i, th0, None, None, use_oracle_obs=False)
        assert abs(ll - rO["ll"][i]) <= 1e-8 * abs(rO["ll"][i])
        assert abs(np.exp(th[1]) - pO.mat_mu[i, 0]) <= 1e-5 * pO.mat_mu[i, 0]


def test_exact_sum_is_order_independent(hh):
    """lp_exact_add: the fixed-point words (and hence the value) do not depend on the summation order,
    which is what makes an N-shard run bit-identical to the unsharded run."""
    import ctypes as C
    rng = np.random.default_rng(0)
    hh.hh_exact_sum.restype = C.c_double
    v = np.ascontiguousarray(-rng.exponential(50.0, 5000) - rng.uniform(0, 1e-3, 5000))
    w0 = (C.c_longlong * 3)()
    r0 = hh.hh_exact_sum(v.ctypes.data_as(A.c_double_p), len(v), w0)
    for _ in range(5):
        p = np.ascontiguousarray(rng.permutation(v))
        w = (C.c_longlong * 3)()
        r = hh.hh_exact_sum(p.ctypes.data_as(A.c_double_p), len(p), w)
        assert list(w) == list(w0) and r == r0
    import math
    assert abs(r0 - math.fsum(v)) <= 1e-15 * abs(math.fsum(v))     # and it is the correctly accumulated sum
    bad = np.array([1.0, np.nan, 2.0])
    wb = (C.c_longlong * 3)()
    assert np.isnan(hh.hh_exact_sum(bad.ctypes.data_as(A.c_double_p), 3, wb)) and wb[2] == 1
