"""Profiling driver: 1M pair, a few stand-alone NN batches and one 50-iteration inner loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
d = synth.make_pair(n, with_clouds=False)
ctx = P.Context(0)
if len(sys.argv) > 3:
    ctx.set_cells_per_point(float(sys.argv[3]))
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
def morton_order(p, h):
    c = np.floor((p - p.min(0)) / h).astype(np.uint64)
    def spread(v):
        v = v & np.uint64(0x1fffff)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1f00000000ffff)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return np.argsort(spread(c[:, 0]) | (spread(c[:, 1]) << np.uint64(1)) | (spread(c[:, 2]) << np.uint64(2)), kind="stable")
for _ in range(2):
    ctx.nn(d["ct2"]); print("nn ms (caller order)", ctx.last_device_ms())
qs = d["ct2"][morton_order(d["ct2"], 0.1)]
for _ in range(2):
    ctx.nn(qs); print("nn ms (morton order)", ctx.last_device_ms())
for _ in range(2):
    r = ctx.icp_run(P.icp_params(max_iter=iters, force_iters=1)); print("icp ms", r["device_ms"], r["grid_blocks"])
