"""GPU: the CUDA path against outputs of the reference's own R source (tests/golden/reference_exec.json, made by
executing /root/reference/R/*.R with tests/r_eval.py -- see tests/test_reference_exec.py for the CPU side, where the
oracle reproduces the same outputs bit for bit).  Tolerances are the north_star's: log-likelihoods 1e-6 relative,
parameters 1e-4 where the optimum is well conditioned; impulse fits statistically (three-start BFGS on 26 cells)."""
import numpy as np
import pytest

from lineagepulse_b200 import _cabi as A
from tests.helpers import rel
from tests.test_reference_exec import FITS, _design

pytestmark = pytest.mark.gpu

NAMES = ("H0 constant ZINB, drop-out EM (fit A)", "H1 impulse ZINB given fit A's drop-out model (fit B), lm by numpy QR",
         "constant NB (fit C)", "groups + batch confounder ZINB (C4 shape), size factors", "splines (df 4) ZINB (C5 shape), size factors",
         "dispersion groups + dispersion batch confounder, NB", "H0 constant ZINB, logistic_ofMu drop-out EM")


def test_gpu_fits_against_the_executed_reference():
    from lineagepulse_b200.engine import Engine
    engines = {}
    n = 0
    for case in FITS:
        if case["name"] not in NAMES:
            continue
        X, design, m = _design(case)
        key = X.tobytes()
        if key not in engines:
            engines[key] = Engine(0)
            engines[key].load_counts_i32(X)
        eng = engines[key]
        eng.set_design(design)
        how, ref = case["how"], case["reference"]
        fit_drop = how["strDropModel"] != "none" and "matDropoutLinModel" not in how
        p = A.ParamsData(design, m)
        if "matDropoutLinModel" in how:
            p.mat_drop = np.asfortranarray(np.asarray(how["matDropoutLinModel"], dtype=np.float64))
        r = eng.fit_model(m, p, fit_drop)
        ll = eng.eval_loglik(m, p)
        ll_ref = np.asarray(ref["vecLL"])
        em = np.asarray(ref["vecEMLogLikModel"])
        assert r["n_iter"] == len(em), case["name"]
        if how["strMuModel"] == "impulse":
            assert np.mean(rel(ll, ll_ref) < 1e-6) >= 0.3 and rel(ll, ll_ref).max() < 5e-3, (case["name"], rel(ll, ll_ref))
        else:
            assert rel(ll, ll_ref).max() < 1e-6, (case["name"], rel(ll, ll_ref))
            assert rel(r["em_ll"][: len(em)], em).max() < 1e-7, case["name"]
            mu_ref = np.asarray(ref["matMuModel"], dtype=np.float64).reshape(p.mat_mu.shape)
            if how["strMuModel"] != "splines":
                assert rel(p.mat_mu, mu_ref).max() < 1e-3, case["name"]
        n += 1
    assert n == len(NAMES)
