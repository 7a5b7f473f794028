"""Grid build time at several sizes, for A/B runs of several builds (PWICP_LIB=<path>).  python scripts/build_probe.py [n ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import pwicp_b200 as P
from pwicp_b200 import synth
if os.environ.get("PWICP_LIB"):
    P._lib = P.load_library(os.environ["PWICP_LIB"])
ctx = P.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1000000, 10000000]:
    d = synth.make_pair(n, with_clouds=False)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ts = []
    for _ in range(8):
        ctx.flush_l2(); ts.append(ctx.target_rebuild())
    print(f"{os.environ.get('PWICP_LIB', 'main'):>40s}  n={len(d['ct1'])}: build min {min(ts):.3f} ms, median {sorted(ts)[4]:.3f} ms", flush=True)
