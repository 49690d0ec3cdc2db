"""Run on a B200: every (gene, start) work item of fitMuDisp on BASELINE config C1 (200 x 200, seed 1) as the CUDA
kernel computes it -> gpurun_out/c1_fit_items.npz (copied to tests/golden/ and compared with the host mirror bit for
bit by tests/test_mirror.py on any box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200.engine import Engine
from tests.helpers import make_dataset, design_for, model
from tests.oracle_binding import Oracle

c1 = make_dataset(200, 70, 65, 65, seed=1)
d = design_for(c1)
o = Oracle()
e = Engine(0)
e.load_counts_i32(c1["X"])
e.set_design(d)
out = {}
for mu in ("impulse", "constant"):
    for drop in ("logistic", "none"):
        m = model(mu, drop=drop)
        p0 = o.init_params(c1["X"], d, m)
        if drop != "none":
            p0.mat_drop[:, 0] = -2.2
        r = e.fit_mudisp_items(m, p0)
        tag = f"{mu}_{drop}"
        out[tag + "_theta0"], out[tag + "_theta"], out[tag + "_ll"] = r["theta0"], r["theta"], r["ll_items"]
        out[tag + "_conv"], out[tag + "_evals"], out[tag + "_n_starts"] = r["conv_items"], r["evals_items"], np.int32(r["n_starts"])
        print(tag, "items", len(r["ll_items"]), "evals", int(r["evals_items"].sum()))
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/c1_fit_items.npz", **out)
