"""Where do GPU and oracle differ at dispersions near the 1e10 clamp?  Per-observation comparison through
evalLogLikMatrix on 1-cell genes: NB / ZINB x zero / non-zero observation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine
from tests.oracle_binding import Oracle

rng = np.random.default_rng(0)
G = 4000
o = Oracle()
for phi in (1e10, 1e8, 1e6, 1e3):
    for kind in ("zero", "nonzero"):
        mu = np.exp(rng.uniform(-3, 7, G))
        x = np.zeros((G, 1), dtype=np.int32) if kind == "zero" else np.maximum(1, rng.poisson(np.minimum(mu, 1e4))).astype(np.int32).reshape(G, 1)
        d = A.DesignData(G, 1)
        e = Engine(0); e.load_counts_i32(x); e.set_design(d)
        for drop in (A.DROP_NONE, A.DROP_LOGISTIC):
            m = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, drop, 0)
            p = A.ParamsData(d, m)
            p.mat_mu[:, 0] = mu; p.mat_disp[:, 0] = phi
            if drop != A.DROP_NONE: p.mat_drop[:, 0] = -2.0
            g, c = e.eval_loglik(m, p), o.eval_loglik(x, d, m, p)
            r = np.abs(g - c) / np.abs(c)
            # what one ulp of log(phi) is worth in the NB zero term phi (log phi - log(phi + mu))
            ulp_bound = phi * 2 * np.spacing(np.log(phi)) / np.abs(c)
            print(f"phi {phi:g} {kind:8s} drop {drop}: max rel {r.max():.3e} median {np.median(r):.3e}  frac identical {(g == c).mean():.3f}  max(rel / two-ulp-of-log bound) {np.max(r / ulp_bound):.3f}")
