"""Host / device time of fit A (constant ZINB + drop-out EM) alone: LPB200_DEBUG=1 python scripts/profile_fitA.py [genes] [cells]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine
from lineagepulse_b200.splines import ns
import bench
genes = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
X, t = bench.make_workload(genes, cells, seed=2)
d = A.DesignData(X.shape[0], X.shape[1], pseudotime=t, impulse_basis=ns(t, 4))
e = Engine(0); e.load_counts_i32(X); e.set_design(d)
mA = A.Model(A.MU_CONSTANT, A.DISP_CONSTANT, A.DROP_LOGISTIC, A.DROPFIT_PERCELL)
for rep in range(3):
    p = A.ParamsData(d, mA); e.reset_stats(); t0 = time.time(); r = e.fit_model(mA, p, True)
    print("fit A wall ms", (time.time() - t0) * 1e3, "cycles", r["n_iter"], e.stats(), flush=True)
