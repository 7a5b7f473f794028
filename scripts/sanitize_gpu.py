"""Small device workload for compute-sanitizer (memcheck / racecheck): the F4 pre-processing kernels, patch statistics, the
stand-alone searches and one short inner loop on small inputs.
    compute-sanitizer --tool memcheck python scripts/sanitize_gpu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
ctx = P.Context(0)
c = synth.make_scan(extent=0.5, spacing=0.005, seed=1)
v = ctx.voxel_grid(c, 0.01); md = ctx.knn_mean_dist(c, 14); p = ctx.preprocess(c, 0.005)
d = synth.make_pair(3000)
ctx.upload_pair(d)
ps = ctx.patch_stats(d["patch_pts2"], d["patch_off2"])
idx, d2 = ctx.nn(d["bp2"]); s = ctx.self_nn(d["ct1"]); q = ctx.percentile_nn(d["cloud1"], d["cloud2"]); o = ctx.overlap_ratio(d["cloud1"], d["cloud2"], 0.05)
ctx.icp_source_upload(d["ct2"])
r = ctx.icp_run(P.icp_params(max_iter=6, force_iters=1))
# the same loop as two launches with the stand-alone search of iteration 1 in between (the path of large source sets),
# the host-buffer call (normals uploaded last) and the outer loop
os.environ["PWICP_SPLIT_MIN_POINTS"] = "0"
r2 = ctx.icp_run(P.icp_params(max_iter=6, force_iters=1))
assert np.array_equal(r["T"], r2["T"]), "split launch changed the result"
h = ctx.icp_p2plane(d["ct1"], d["nrm1"], d["ct2"], P.icp_params(max_iter=6, force_iters=1))
del os.environ["PWICP_SPLIT_MIN_POINTS"]
ctx.upload_pair(d)
g = ctx.piecewise_icp(P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"]), 1, 0.05)
print("sanitize workload ok:", len(v), len(p), len(idx), r["n_iter"] if "n_iter" in r else "")
