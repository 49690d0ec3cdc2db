import sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from lineagepulse_b200 import api
X, t = bench.make_workload(5000, 10000, seed=2)
G, N = X.shape
pinned = torch.empty((G, N), dtype=torch.int32).pin_memory(); pinned.numpy()[...] = X
annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
api.runLineagePulse(pinned.numpy(), annot, strMuModel="impulse", strDropModel="logistic", boolVerbose=False)
pr = cProfile.Profile(); t0 = time.time(); pr.enable()
obj = api.runLineagePulse(pinned.numpy(), annot, strMuModel="impulse", strDropModel="logistic", boolVerbose=False)
pr.disable(); print("wall", time.time() - t0)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
