"""Search-side probe: the stand-alone correspondence consumers (pwicp_nn, the outer loop's classification / percentile
kernels) and the inner loop's pre-pass, for A/B runs of several builds (PWICP_LIB=<path>).
python scripts/search_probe.py [n_patches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
from oracle import oracle_py as O
if os.environ.get("PWICP_LIB"):
    P._lib = P.load_library(os.environ["PWICP_LIB"]); print("library", os.environ["PWICP_LIB"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
d = synth.make_pair(n)
ctx = P.Context(0)
ctx.upload_pair(d)
# (a) unseeded NN, caller order (rows of the generator grid) and a random order
q = d["bp2"]
for name, qq in (("caller order", q), ("random order", q[np.random.default_rng(0).permutation(len(q))])):
    ts = []
    for _ in range(3):
        idx, d2 = ctx.nn(qq); ts.append(ctx.last_device_ms())
    print(f"pwicp_nn {len(qq)} queries, {name}: {min(ts):.3f} ms")
oi, od = O.nn(d["ct1"], q[:200000])
idx, d2 = ctx.nn(q[:200000])
print("  parity vs oracle (200k): idx", int((idx != oi).sum()), "d2", int((d2 != od).sum()))
# (b) outer loop
pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
for rep in range(3):
    ctx.upload_pair(d)
    g = ctx.piecewise_icp(pp, 1, 0.05)
print("outer loop:", g["n_outer"], "iterations, device ms", round(g["device_ms"], 3), [round(s.device_ms, 3) for s in g["stats"]])
ref = O.piecewise_icp(O.PairData(d), 1, 0.05)
print("  DTseries equal", np.array_equal(g["DTseries"], ref["DTseries"]), "n_stable equal",
      [a.n_stable for a in g["stats"]] == [b.n_stable for b in ref["stats"]], "max|dT|", float(np.abs(g["T"] - ref["T"]).max()))
# (c) inner loop at the same size: sort + pre-pass = loop - kernel
ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"]); ctx.icp_source_upload(d["ct2"])
best = min((ctx.icp_run(P.icp_params(max_iter=10, force_iters=1)) for _ in range(4)), key=lambda r: r["device_ms"])
print(f"inner loop 10 iterations: {best['device_ms']:.3f} ms, kernel {best['kernel_ms']:.3f} ms, sort + pre-pass {best['device_ms'] - best['kernel_ms']:.3f} ms")
