"""Dumps the hardware reciprocal seed MUFU.RCP64H (rcp.approx.ftz.f64) on the GPU box: all 2^20 upper-word
mantissas at exponent 0, plus probes of the exponent / lower-word (in)dependence.  Output:
gpurun_out/rcp64h_raw.npz, from which scripts/make_rcp64h_golden.py builds tests/golden/rcp64h.npz."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lineagepulse_b200.engine import Engine

e = Engine(0)
m = np.arange(1 << 20, dtype=np.int64)
out = {}
for name, expo in (("e3ff", 0x3FF), ("e400", 0x400), ("e3fe", 0x3FE), ("e001", 0x001), ("e7fd", 0x7FD), ("e200", 0x200)):
    hi = ((expo << 20) | m).astype(np.int32)
    oh, ol = e.rcp64h_probe(hi, 0)
    out[name + "_hi"], out[name + "_lo"] = oh, ol
hi = ((0x3FF << 20) | m).astype(np.int32)
oh, ol = e.rcp64h_probe(hi, -1)          # lower word all ones: is it ignored?
out["e3ff_loFF_hi"], out["e3ff_loFF_lo"] = oh, ol
hi = ((0xBFF << 20) | m).astype(np.int32)  # negative arguments
oh, ol = e.rcp64h_probe(hi, 0)
out["neg_hi"], out["neg_lo"] = oh, ol
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/rcp64h_raw.npz", **out)
print("lo ignored:", np.array_equal(out["e3ff_hi"], out["e3ff_loFF_hi"]), "lo words all zero:", all(not out[k].any() for k in out if k.endswith("_lo")))
base = out["e3ff_hi"].astype(np.int64)
for name, expo in (("e400", 0x400), ("e3fe", 0x3FE), ("e001", 0x001), ("e7fd", 0x7FD), ("e200", 0x200)):
    d = out[name + "_hi"].astype(np.int64) - base
    print(name, "hi-word offset vs exponent 0: unique", np.unique(d)[:8], "expected", (0x3FF - expo) << 20)
