"""The north_star sentence literally: ONE runLineagePulse call on the 20 000 x 100 000 matrix (the matrix of bench.py's
strong pass: seed 3) from host counts through the public API, on n devices of this process.

    python scripts/c3_public_api.py [n_devices] [genes] [cells]

Prints the wall time of the call, the DE count and the digest bench.py's strong pass prints for the same matrix (per-gene
log-likelihoods of the four fits, p, padj): equal digests = the public API on n devices returns the sharded ranks' bits."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lineagepulse_b200 import api
from lineagepulse_b200.simulate import simulate_rows_device, workload_recipe

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
G = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
rec = workload_recipe(G, seed=3)
pinned = torch.empty((G, N), dtype=torch.int32).pin_memory()
t0 = time.time()
for lo in range(0, G, 2000):      # generated on the device in slabs, parked in pinned host memory
    hi = min(G, lo + 2000)
    pinned[lo:hi].copy_(simulate_rows_device(rec, N, lo, hi, "cuda:0"))
torch.cuda.synchronize()
torch.cuda.empty_cache()
t_gen = time.time() - t0
t = np.arange(N, dtype=np.float64) / max(N - 1, 1)
annot = {"cell": np.array([f"cell_{j + 1}" for j in range(N)]), "continuous": t}
kw = dict(strMuModel="impulse", strDropModel="logistic", boolVerbose=False, devices=list(range(ndev)) if ndev > 1 else None)
small = api.runLineagePulse(pinned.numpy()[:64 * ndev], annot, **kw)       # first-use costs (communicator, modules) outside the timing
small.matCountsProc.engine().close()
t0 = time.time()
obj = api.runLineagePulse(pinned.numpy(), annot, **kw)
dt = time.time() - t0
res = obj.dfResults
ll = np.stack([res[c].values for c in ("loglik_full_zinb", "loglik_red_zinb", "loglik_full_nb", "loglik_red_nb")], 1)
h = hashlib.sha256()
for a in (ll, res["p"].values, res["padj"].values):
    h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
print(json.dumps({"workload": f"C3 {G}x{N} through runLineagePulse(host counts, devices={kw['devices']})", "n_devices": ndev,
                  "call_s": dt, "genes_per_s": G / dt, "de_genes_q05": int((res["padj"] < 0.05).sum()),
                  "all_zero_genes": int(res["allZero"].sum()), "digest": h.hexdigest()[:32], "generation_s": t_gen,
                  "h2d_bytes": int(pinned.numel() * 4)}))
