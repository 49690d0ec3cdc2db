"""Profiling driver: one pass of fits A-D + LRT on a matrix sized for one full wave of CTAs
(default 592 genes x 10 000 cells), used under ncu (see profiles/README.md)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine
from lineagepulse_b200.splines import ns
import bench

genes = int(sys.argv[1]) if len(sys.argv) > 1 else 592
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
X, t = bench.make_workload(genes, cells, seed=2)
d = A.DesignData(X.shape[0], X.shape[1], pseudotime=t, impulse_basis=ns(t, 4))
e = Engine(0)
e.load_counts_i32(X)
e.set_design(d)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for _ in range(reps):
    t0 = time.time()
    res, work = bench.native_pipeline(e, d, A, float("nan"))
    print("pipeline s", time.time() - t0, {k: (v["ole"], round(v["kernel_ms"], 1)) for k, v in work.items()})
