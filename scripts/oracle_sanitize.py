"""Runs every entry point of the CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer.

    make -C oracle sanitize        # builds oracle/_san/liboracle_san.so and runs this script with the runtimes preloaded
"""
import sys, os, ctypes as C
ROOT = os.environ.get("PWICP_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
from oracle import oracle_py as O
# swap in the sanitised build
O._LIB = None
_orig = C.CDLL
def patched(path, *a, **k):
    if str(path).endswith("liboracle.so"): path = os.path.join(ROOT, "oracle", "_san", "liboracle_san.so")
    return _orig(path, *a, **k)
C.CDLL = patched
O.C.CDLL = patched
from pwicp_b200 import synth
d = synth.make_pair(3000)
idx, d2 = O.nn(d["ct1"], np.concatenate([d["ct2"], d["bp2"]]))
r = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=8, force_iters=1), trace=True)
r1 = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=8, force_iters=1, reduce_mode=1, group_batches=32))
res = O.piecewise_icp(O.PairData(d), 1, 0.05)
res2 = O.piecewise_icp(O.PairData(d), 0, 0.0)
V, s = O.vcm(d["ct1"], d["nrm1"], d["ct2"][:500])
ps = O.patch_stats(d["patch_pts2"], d["patch_off2"])
c = synth.make_scan(extent=0.6, spacing=0.005, seed=1)
p = O.preprocess(c, 0.005, 14, 5.0)
print("sanitised oracle run ok:", len(idx), r["n_iter"], len(res["DTseries"]), len(res2["DTseries"]), ps["ct"].shape, p.shape)
