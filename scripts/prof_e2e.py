"""Where the time of the public call goes: python scripts/prof_e2e.py [n_devices] [genes_per_device] [cells]."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from lineagepulse_b200 import api
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
genes = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
X = np.concatenate([bench.make_workload(genes, cells, seed=2 + 1000 * r)[0] for r in range(ndev)], 0)
t = np.arange(cells) / (cells - 1)
G, N = X.shape
pinned = torch.empty((G, N), dtype=torch.int32).pin_memory(); pinned.numpy()[...] = X
annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
kw = dict(strMuModel="impulse", strDropModel="logistic", boolVerbose=False, devices=list(range(ndev)) if ndev > 1 else None)
t0 = time.time(); api.runLineagePulse(pinned.numpy(), annot, **kw); print("first call wall", time.time() - t0, flush=True)
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
obj = api.runLineagePulse(pinned.numpy(), annot, **kw)
pr.disable(); print("second call wall", time.time() - t0, "genes/s", G / (time.time() - t0))
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:7000])
