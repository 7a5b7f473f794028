"""Generates tests/golden/refpair_e2.npz: BASELINE configs[0], the reference's shipped two-epoch pair
(data/data_synthetic/syntheticPC_with_transformations/Epoch_001.pcd -> Epoch_002.pcd) at the hot-path boundary,
together with the result the reference recorded for it (results/4DPCReg/2_Direct2Ref_TransMatrix.txt).

    python tests/golden/make_refpair.py          (needs /root/reference and oracle/_ref/libref_supervoxel.so)

Inputs stored: the two pre-processed, centroid-shifted clouds (what Piecewise_ICP receives, src/Registration.cpp:452)
and, per cloud point, the index of the selected patch it belongs to (-1: none) as produced by the REFERENCE'S OWN
supervoxel segmentation (codelibrary, compiled where it lies: oracle/ref_supervoxel.cpp) followed by the reference's
patch refinement / planarity gates (host mirror).  Patch k = the cloud points labelled k, in cloud order
(src/Segmentation.cpp:95-100), so the centroid-level pair is rebuilt from these arrays by refpair.load().
Outputs stored: the reference's recorded 4x4 and 6x6 VCM (12 printed decimals) and the ground-truth matrix.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import refdata_oracle as R                  # noqa: E402
from pwicp_b200 import host                 # noqa: E402
import ctypes as C                          # noqa: E402


def labels_of(cloud, patch, off):
    """index of the patch every cloud point belongs to (-1: none); patch points are exact copies of cloud points"""
    key = lambda a: np.ascontiguousarray(a, np.float32).view([("", np.float32)] * 3).ravel()
    ck, pk = key(cloud), key(patch)
    order = np.argsort(ck, kind="stable")
    pos = np.searchsorted(ck[order], pk)
    idx = order[pos]
    assert np.array_equal(ck[idx], pk) and len(np.unique(idx)) == len(idx), "duplicate coordinates: labels ambiguous"
    lab = np.full(len(cloud), -1, np.int16)
    lab[idx] = np.repeat(np.arange(len(off) - 1), np.diff(off)).astype(np.int16)
    for k in (0, len(off) // 2, len(off) - 2):                 # cloud order within a patch is preserved
        assert np.array_equal(cloud[lab == k], patch[off[k]:off[k + 1]])
    return lab


def main(epoch=2):
    scans = os.path.join(R.REF, "data/data_synthetic/syntheticPC_with_transformations")
    e1 = host.load_pcd(os.path.join(scans, "Epoch_001.pcd"))
    e2 = host.load_pcd(os.path.join(scans, "Epoch_%03d.pcd" % epoch))
    L = host.lib()
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_supervoxel.so"))
    L.pwicp_host_set_segmenter(C.cast(ref.ref_supervoxel_labels, C.c_void_p))
    pp = R.prepare_pair(e1, e2)
    L.pwicp_host_set_segmenter(None)
    Tr, Vr = R.read_T(os.path.join(R.REF, "results/4DPCReg/%d_Direct2Ref_TransMatrix.txt" % epoch))
    gt = R.ground_truth(os.path.join(R.REF, "data/data_synthetic/defined_transformations.txt"))[epoch]
    out = os.path.join(HERE, "refpair_e%d.npz" % epoch)
    np.savez_compressed(out, cloud1=pp["cloud1"], cloud2=pp["cloud2"],
                        lab1=labels_of(pp["cloud1"], pp["patch1"], pp["off1"]),
                        lab2=labels_of(pp["cloud2"], pp["patch2"], pp["off2"]),
                        shift=pp["shift"], T_recorded=Tr, VCM_recorded=Vr, T_truth=gt,
                        config=np.array([R.RES, R.SV, R.DTINIT, R.DTMIN], np.float32))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
