"""Tiny end-to-end pass for compute-sanitizer (memcheck / racecheck): every kernel runs at least once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine, lrt
from lineagepulse_b200.simulate import simulateContinuousDataSet
from lineagepulse_b200.splines import ns

sim = simulateContinuousDataSet(72, 6, 5, 5, scaMumax=60, scaSDMuAmplitude=2, vecGeneWiseDropoutRates=np.full(16, 0.1), seed=3)
X = np.ascontiguousarray(sim["counts"][sim["counts"].sum(1) > 0], dtype=np.int32)
X[1, 3] = -1
t = sim["annot"]["continuous"]
G, N = X.shape
rng = np.random.default_rng(0)
groups = np.repeat(np.arange(3), N // 3)
batch = rng.integers(0, 2, N)
pred = np.log(np.maximum(X, 0).mean(1) + 1).reshape(-1, 1)
d = A.DesignData(G, N, size_factors=rng.uniform(0.7, 1.4, N), pseudotime=t, groups=groups, spline_basis_mu=ns(t, 4),
                 impulse_basis=ns(t, 4), confounders_mu=[batch], pi_const_pred=pred)
e = Engine(0)
e.load_counts_dense(np.where(X < 0, np.nan, X).astype(np.float64))
e.set_design(d)
ctl = A.default_control(); ctl.maxit_mudisp = 25; ctl.maxit_pi = 25
def M(mu, drop, grp=0): return A.Model(A.MU_MODELS[mu], A.DISP_CONSTANT, A.DROP_MODELS[drop], grp)
pA = A.ParamsData(d, M("constant", "logistic")); e.fit_model(M("constant", "logistic"), pA, True, ctl)
jobs = []
for mu, drop in (("impulse", "logistic"), ("groups", "none"), ("splines", "none")):
    p = A.ParamsData(d, M(mu, drop))
    if drop != "none": p.mat_drop = pA.mat_drop.copy()
    jobs.append((M(mu, drop), p))
res = e.fit_model_multi(jobs, ctl)
pO = A.ParamsData(d, M("constant", "logistic_ofMu")); e.fit_model(M("constant", "logistic_ofMu"), pO, True, ctl)
pF = e.init_params(M("constant", "logistic", 1)); e.fit_pi(M("constant", "logistic", 1), pF, ctl)
ll = e.eval_loglik(jobs[0][0], jobs[0][1])
for what in (A.EXPAND_MEAN, A.EXPAND_DISP, A.EXPAND_DROP, A.EXPAND_POSTDROP, A.EXPAND_NORMDATA):
    e.expand_fits(jobs[0][0], jobs[0][1], [0, G - 1], what)
print("sanitize pass done", np.isfinite(ll).all(), [r["ll"].sum() for r in res])
