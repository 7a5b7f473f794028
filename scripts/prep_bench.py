"""F4 timing: PCpreprocessing (VoxelGrid + StatisticalOutlierRemoval, k = 14) on the device against the CPU restatement
(oracle, single thread) on the same clouds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
from oracle import oracle_py as O
ctx = P.Context(0)
for name, c, leaf in (("scan 2 m / 5 mm", synth.make_scan(extent=2.0, spacing=0.005, seed=9), 0.005),
                      ("1M centroids", synth.make_pair(1000000, with_clouds=False)["ct1"], 0.05),
                      ("4M surface", synth.make_scan(extent=10.0, spacing=0.005, seed=2), 0.005)):
    for _ in range(2):
        ctx.voxel_grid(c, leaf); tv = ctx.last_device_ms()
        ctx.knn_mean_dist(c, 14); tk = ctx.last_device_ms(); tkk = ctx.last_knn_kernel_ms()
        t0 = time.time(); out = ctx.preprocess(c, leaf); wall = (time.time() - t0) * 1e3; tp = ctx.last_device_ms()
    t0 = time.time(); ov = O.voxel_grid(c, leaf); cv = (time.time() - t0) * 1e3
    t0 = time.time(); O.knn_mean_dist(c, 14); ck = (time.time() - t0) * 1e3
    print("%-16s n=%8d -> %8d | device: voxel grid %.3f ms, 14-NN mean distance %.3f ms incl. grid build, kernel alone %.3f ms (%.1f M pts/s), PCpreprocessing %.3f ms device / %.1f ms wall incl. copies | CPU: voxel grid %.0f ms, 14-NN %.0f ms"
          % (name, len(c), len(out), tv, tk, tkk, len(c) / tkk / 1e3, tp, wall, cv, ck), flush=True)
