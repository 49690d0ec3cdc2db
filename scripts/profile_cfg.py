"""Profiling driver for the other BASELINE configurations: one pass of fits A-D + LRT of a C4- or C5-shaped matrix
(counts generated on the device), used under ncu:  python scripts/profile_cfg.py c4|c5 [genes] [cells]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200 import _cabi as A
from lineagepulse_b200.engine import Engine
from lineagepulse_b200.simulate import simulate_rows_device, workload_recipe
from lineagepulse_b200.splines import ns
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
G = int(sys.argv[2]) if len(sys.argv) > 2 else 592
N = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
t, groups, batch = bench.cell_covariates(N, seed=4)
if cfg == "c4":
    rows = simulate_rows_device(workload_recipe(G, seed=4, n_groups=4, n_batches=3), N, 0, G, "cuda:0", groups, batch)
    d = A.DesignData(G, N, pseudotime=t, groups=groups, confounders_mu=[batch])
    h1 = "groups"
else:
    rows = simulate_rows_device(workload_recipe(G, seed=5), N, 0, G, "cuda:0")
    d = A.DesignData(G, N, pseudotime=t, spline_basis_mu=ns(t, 6))
    h1 = "splines"
e = Engine(0)
e.load_counts_device_i32(rows.data_ptr(), G, N)
e.set_design(d)
t0 = time.time()
res, work = bench.native_pipeline(e, d, A, float("nan"), h1)
print(cfg, "pipeline s", time.time() - t0, {k: (v["ole"], round(v["kernel_ms"], 1)) for k, v in work.items()})
