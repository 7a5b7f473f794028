"""Lean A/B harness (1M pair only): bench-shaped 50-iteration loop and the steady-state cost per iteration, one process per
library so that every build sees the same box.   usage: python scripts/ab_lean.py libA.so libB.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
P._lib = P.load_library(sys.argv[1])
ctx = P.Context(0)
d = synth.make_pair(int(os.environ.get("AB_N", "1000000")), with_clouds=False)
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
ctx.icp_source_upload(d["ct2"])
for _ in range(6):
    ctx.icp_run(P.icp_params(max_iter=300, force_iters=1))
t = {}
for it in (1, 2, 3, 4, 5, 50, 250, 450):
    rs = [ctx.icp_run(P.icp_params(max_iter=it, force_iters=1)) for _ in range(5)]
    t[it] = (min(r["device_ms"] for r in rs), min(r["kernel_ms"] for r in rs))
r = ctx.icp_run(P.icp_params(max_iter=50, force_iters=1))
import zlib
print("%%-22s n %%d grid %%4d x %%d warps | it1 %%.3f (+%%.3f +%%.3f +%%.3f +%%.3f)  it50 %%.3f ms (kernel %%.3f) | steady %%.2f us/iter | T crc %%08x" %% (
    os.path.basename(sys.argv[1]), len(d["ct2"]), r["grid_blocks"], r["warps_per_block"], t[1][0], t[2][0] - t[1][0], t[3][0] - t[2][0],
    t[4][0] - t[3][0], t[5][0] - t[4][0], t[50][0], t[50][1], (t[450][0] - t[250][0]) / 200 * 1e3, zlib.crc32(r["T"].tobytes())))
''' % ROOT
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(lib)])
