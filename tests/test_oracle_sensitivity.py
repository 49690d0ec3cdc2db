"""The reference algorithm's own numerical conditioning (CPU-only, documents the parity bar).

R's sum() accumulates in 80-bit long double on x86 and in plain double where long double is
unavailable -- both are legitimate R behaviours.  Running the oracle in both modes shows that
well-conditioned fits (constant mean) agree to ~1e-14, while the 3-start impulse fits agree
within 1e-6 on only ~60 % of genes: BFGS with finite-difference gradients amplifies last-digit
differences of the objective.  This is the floor any re-implementation (R on another platform
included) is subject to, and it is the bar the GPU impulse-fit tests use.
"""
import numpy as np

from lineagepulse_b200 import _cabi as A
from tests.helpers import design_for, model, rel


def test_reference_algorithm_self_agreement(oracle, c1):
    X = c1["X"]
    d = design_for(c1)
    res = []
    try:
        for mode in (0, 1):
            oracle.lib.lporacle_set_sum_mode(mode)
            mA = model("constant", drop="logistic")
            pA = A.ParamsData(d, mA)
            rA = oracle.fit_model(X, d, mA, pA, True)
            mB = model("impulse", drop="logistic")
            pB = A.ParamsData(d, mB)
            pB.mat_drop = pA.mat_drop.copy()
            rB = oracle.fit_model(X, d, mB, pB, False)
            res.append((rA["ll"], rB["ll"]))
    finally:
        oracle.lib.lporacle_set_sum_mode(0)
    (a0, b0), (a1, b1) = res
    assert rel(a1, a0).max() < 1e-10              # constant fits: identical
    r = rel(b1, b0)
    frac = (r < 1e-6).mean()
    assert 0.3 < frac < 0.95, frac                # impulse fits: chaotic in the last digits
    assert r.max() > 1e-5
    assert abs(b0.sum() - b1.sum()) <= 1e-4 * abs(b0.sum())
