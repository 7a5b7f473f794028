"""Pins the hot-path restatement (the CPU oracle) against the results the REFERENCE ITSELF recorded.

CPU only.  For every pair of the reference's shipped synthetic series (data/data_synthetic, 20 epochs) this runs
  host mirror  : PCpreprocessing + centroid shift + patch post-processing   (libpwicp_host.so, no device needed)
  oracle/_ref  : the reference's OWN supervoxel segmentation (codelibrary, compiled where it lies; registered through the
                 product's segmenter plug-in pwicp_host_set_segmenter)
  oracle       : Piecewise_ICP outer loop (classification, DT schedule, inner ICP, VCM)
with the shipped configuration (configuration_files/configuration_4d.txt) and compares the final 4x4 with
results/4DPCReg/<epoch>_Direct2Ref_TransMatrix.txt and with the ground truth (defined_transformations.txt).

    python scripts/refdata_oracle.py [first_epoch last_epoch] [--standin] [--mode direct|fixed|adaptive]

--msvc-order: pcl::VoxelGrid sums the points of a voxel in the order its (unstable) std::sort leaves them; the recorded results
come from the reference's Windows build, i.e. the Microsoft STL's order (host/msvc_sort.h) -- with it every recorded pair is
reproduced.  --standin: the library's cubic-cell stand-in instead of oracle/_ref.

--mode fixed: the recorded <e>_Fixed_TransMatrix.txt files (interval 3: epoch e against epoch e - 3, the reference epoch
for e <= 4; the interval is not recorded, 3 is the one that reproduces the files).  --mode adaptive: the recorded
<e>_Adaptive_TransMatrix.txt files with the pairs calAdaptivePairSequence selects for this series (overlap threshold 0.75,
DTinit 0.05; RegPairFile.txt of the device run, gpurun_out/refdata_4d/Adaptive).

Remaining differences to the recorded numbers come from the pre-processing in front of the segmentation, which is PCL in
the reference (VoxelGrid / StatisticalOutlierRemoval; summation order inside a voxel is unspecified) and a host mirror here.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
from oracle import oracle_py as O          # noqa: E402
from pwicp_b200 import host                # noqa: E402

REF = "/root/reference"
RES, SV, DTINIT, DTMIN = 0.005, 0.05, 0.05, 0.004        # configuration_files/configuration_4d.txt:4-10


def prepare_pair(xyz1, xyz2, res=RES, sv=SV):
    L = host.lib()
    L.pwicp_host_prepare_pair.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_float] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    a, b = np.ascontiguousarray(xyz1, np.float32), np.ascontiguousarray(xyz2, np.float32)
    cap = max(len(a), len(b))
    c1 = np.zeros((cap, 3), np.float32); c2 = np.zeros((cap, 3), np.float32)
    p1 = np.zeros((cap, 3), np.float32); p2 = np.zeros((cap, 3), np.float32)
    o1 = np.zeros(cap + 1, np.int32); o2 = np.zeros(cap + 1, np.int32)
    shift = np.zeros(3, np.float32); sz = np.zeros(8, np.int32)
    rc = L.pwicp_host_prepare_pair(a.ctypes.data, len(a), b.ctypes.data, len(b), res, res, sv, sv, c1.ctypes.data, c2.ctypes.data, cap,
                                   p1.ctypes.data, o1.ctypes.data, p2.ctypes.data, o2.ctypes.data, cap, cap, shift.ctypes.data, sz.ctypes.data)
    assert rc == 0
    m1, m2, n1, n2, t1, t2 = sz[:6]
    return {"cloud1": c1[:m1], "cloud2": c2[:m2], "patch1": p1[:t1], "off1": o1[:n1 + 1], "patch2": p2[:t2], "off2": o2[:n2 + 1], "shift": shift}


def centroid_level_pair(pp, res=RES, sv=SV, dtmin=DTMIN):
    """What Piecewise_ICP holds after PatchGenerationAndRefinement + calBPandCTSTD (src/Registration.cpp:653-664)."""
    s1, s2 = O.patch_stats(pp["patch1"], pp["off1"]), O.patch_stats(pp["patch2"], pp["off2"])
    nrm1 = s1["nrm"].copy()
    small = ~((np.diff(pp["off1"]) > 6) & (s1["nrm_ok"] != 0))           # generateCentroidCloudWithPatchNormals :367
    nrm1[small] = (0, 0, 1)
    return {"cloud1": pp["cloud1"], "ct1": s1["ct"], "nrm1": nrm1, "nrm1_ok": s1["nrm_ok"], "ctstd1": s1["ctstd"],
            "cloud2": pp["cloud2"], "ct2": s2["ct"], "bp2": s2["bp"].reshape(-1, 3), "bpstd2": s2["bpstd"],
            "patch_off2": pp["off2"], "patch_pts2": pp["patch2"],
            "Res1": res, "Res2": res, "SVRes1": sv, "SVRes2": sv, "DTmin": dtmin}


def register(xyz1, xyz2, use_reference_segmenter=True):
    L = host.lib()
    L.pwicp_host_set_segmenter.argtypes = [C.c_void_p]
    if use_reference_segmenter:
        ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_supervoxel.so"))
        L.pwicp_host_set_segmenter(C.cast(ref.ref_supervoxel_labels, C.c_void_p))
    try:
        pp = prepare_pair(xyz1, xyz2)
    finally:
        L.pwicp_host_set_segmenter(None)
    d = centroid_level_pair(pp)
    res = O.piecewise_icp(O.PairData(d), 1, DTINIT)
    S = np.eye(4, dtype=np.float32); S[:3, 3] = pp["shift"]
    Si = np.eye(4, dtype=np.float32); Si[:3, 3] = -pp["shift"]
    T = O.mat4_mul(O.mat4_mul(Si, res["T"]), S)                            # src/Registration.cpp:461 (float, Eigen order)
    return T, res, d


def read_T(path):
    l = open(path).read().splitlines()
    return np.array([[float(v) for v in l[1 + r].split()] for r in range(4)]), np.array([[float(v) for v in l[16 + r].split()] for r in range(6)])


def ground_truth(path):
    v = open(path).read().split()
    out, p = {}, 0
    while p < len(v):
        out[int(v[p])] = np.array(v[p + 1:p + 17], float).reshape(4, 4); p += 17
    return out


def pose_err(T, G):
    da = np.abs(O.matrix2angle(np.asarray(T, np.float32)) - O.matrix2angle(np.asarray(G, np.float32)))
    return da.max(), np.abs(np.asarray(T)[:3, 3] - np.asarray(G)[:3, 3]).max()


def main():
    args = [a for a in sys.argv[1:] if a.isdigit()]
    first, last = (int(args[0]), int(args[1])) if len(args) == 2 else (2, 20)
    standin = "--standin" in sys.argv
    scans = os.path.join(REF, "data/data_synthetic/syntheticPC_with_transformations")
    gt = ground_truth(os.path.join(REF, "data/data_synthetic/defined_transformations.txt"))
    mode = sys.argv[sys.argv.index("--mode") + 1] if "--mode" in sys.argv else "direct"
    if "--msvc-order" in sys.argv:
        os.environ["PWICP_VOXEL_ORDER"] = "msvc"     # host/msvc_sort.h: the within-voxel summation order of the reference's Windows build
        print("VoxelGrid: points of a voxel in the order of the Microsoft STL's std::sort")
    tag = {"direct": "Direct2Ref", "fixed": "Fixed", "adaptive": "Adaptive"}[mode]
    load = lambda k: host.load_pcd(os.path.join(scans, "Epoch_%03d.pcd" % k))
    # 1-based target epoch of every source epoch
    adaptive = {2: 1, 3: 1, 4: 1, 5: 1, 6: 1, 7: 3, 8: 4, 9: 4, 10: 5, 11: 6, 12: 6, 13: 7, 14: 9, 15: 12, 16: 13, 17: 14, 18: 14, 19: 14, 20: 14}
    target_of = {"direct": lambda e: 1, "fixed": lambda e: max(1, e - 3), "adaptive": lambda e: adaptive[e]}[mode]
    print("mode %s" % mode)
    print("epoch target | n1 n2 outer | vs recorded: rot[rad] transl[m] VCM rel | vs truth ours: rot transl | vs truth recorded: rot transl | s")
    n_ok = 0
    for e in range(first, last + 1):
        t0 = time.time()
        tgt = target_of(e)
        T, res, d = register(load(tgt), load(e), not standin)
        Tr, Vr = read_T(os.path.join(REF, "results/4DPCReg/%d_%s_TransMatrix.txt" % (e, tag)))
        a, b = pose_err(T, Tr)
        n_ok += int(a <= 1e-6 and b <= 1e-6)
        if tgt == 1:
            G = np.linalg.inv(gt[e])                                     # the files hold reference -> epoch
            g1, g2 = min((pose_err(T, X) for X in (gt[e], G)), key=lambda x: x[0])
            r1, r2 = min((pose_err(Tr, X) for X in (gt[e], G)), key=lambda x: x[0])
        else:
            g1 = g2 = r1 = r2 = float("nan")                             # the truth is given against the reference epoch only
        vrel = np.abs(np.sqrt(np.diag(res["VCM"])) / np.sqrt(np.diag(Vr)) - 1).max()
        print("%5d %6d | %4d %4d %2d | %.2e %.2e %.1e | %.2e %.2e | %.2e %.2e | %.1f" %
              (e, tgt, len(d["ct1"]), len(d["ct2"]), len(res["DTseries"]) - 1, a, b, vrel, g1, g2, r1, r2, time.time() - t0), flush=True)
    print("within 1e-6 rad / 1e-6 m of the recorded matrix: %d of %d" % (n_ok, last - first + 1))


if __name__ == "__main__":
    main()
