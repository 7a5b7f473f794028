"""Sweep of the grid resolution (cells per target point) at 1M: build, unseeded NN, 50-iteration loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
d = synth.make_pair(1000000, with_clouds=False)
ctx = P.Context(0)
for cpp in (1.0, 2.0, 3.0, 4.0, 6.0, 8.0, 12.0):
    ctx.set_cells_per_point(cpp)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])
    b = min(ctx.target_rebuild() for _ in range(3))
    for _ in range(3):
        ctx.icp_run(P.icp_params(max_iter=100, force_iters=1))
    nn = []
    for _ in range(3):
        ctx.nn(d["ct2"]); nn.append(ctx.last_device_ms())
    t = {it: min(ctx.icp_run(P.icp_params(max_iter=it, force_iters=1))["device_ms"] for _ in range(4)) for it in (1, 6, 51, 151)}
    print("cells/pt %5.1f: build %.3f ms, nn %.3f ms, it1 %.3f, it6 %.3f, it51 %.3f, steady %.2f us | step %.3f ms" % (
        cpp, b, min(nn), t[1], t[6], t[51], (t[151] - t[51]) * 10, b + t[51]))
