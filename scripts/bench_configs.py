"""Throughput of the other BASELINE configurations at reduced gene counts (development numbers):
C4-shaped group-wise DE with a batch confounder, C5-shaped splines mean model."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine, lrt
from lineagepulse_b200.splines import ns
import bench

def pipeline(e, d, mu):
    ctl = A.default_control()
    mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mB = A.Model(A.MU_MODELS[mu], A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
    mC = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_NONE, 0)
    mD = A.Model(A.MU_MODELS[mu], A.DISP_CONSTANT, A.DROP_NONE, 0)
    out = {}
    def timed(tag, fn):
        t0 = time.time(); e.reset_stats(); fn(); out[tag] = e.stats(); out[tag]["wall_ms"] = (time.time() - t0) * 1e3
    pA = A.ParamsData(d, mA); timed("A", lambda: e.fit_model(mA, pA, True, ctl))
    pB = A.ParamsData(d, mB); pB.mat_drop = pA.mat_drop.copy(order="F"); timed("B", lambda: e.fit_model(mB, pB, False, ctl))
    pC = A.ParamsData(d, mC); timed("C", lambda: e.fit_model(mC, pC, False, ctl))
    pD = A.ParamsData(d, mD); timed("D", lambda: e.fit_model(mD, pD, False, ctl))
    t0 = time.time()
    llf, llr = e.eval_loglik(mB, pB), e.eval_loglik(mA, pA)
    out["DE"] = {"fn_evals": 0, "kernel_ms": 0.0, "ole": 0, "wall_ms": (time.time() - t0) * 1e3}
    p, padj = lrt(llf, llr, d.deg_freedom(mB) - d.deg_freedom(mA))
    return out, int((padj < 0.05).sum())

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
X, t = bench.make_workload(genes, cells, seed=4)
G, N = X.shape
rng = np.random.default_rng(4)
groups = np.repeat(np.arange(4), (N + 3) // 4)[:N]
batch = rng.integers(0, 3, N)
for name, d, mu in (("C4 groups(4)+batch(3)", A.DesignData(G, N, pseudotime=t, groups=groups, confounders_mu=[batch]), "groups"),
                    ("C5 splines(6)", A.DesignData(G, N, pseudotime=t, spline_basis_mu=ns(t, 6)), "splines")):
    e = Engine(0); e.load_counts_i32(X); e.set_design(d)
    pipeline(e, d, mu)
    t0 = time.time(); w, nde = pipeline(e, d, mu); dt = time.time() - t0
    print(name, f"{G} x {N}: {dt:.2f} s/step, {G / dt:.0f} genes/s, DE {nde};",
          {k: (v["fn_evals"], round(v["kernel_ms"], 1), round(v["wall_ms"], 1), f"{v['ole'] / max(v['kernel_ms'], 1e-9) / 1e6:.1f} GOLE/s") for k, v in w.items()}, flush=True)
