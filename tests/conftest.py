import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "piecewise-icp_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import pwicp_b200
        return pwicp_b200.load_library().pwicp_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def small_pair():
    from pwicp_b200 import synth
    return synth.make_pair(3000)


@pytest.fixture(scope="session")
def gpu_ctx():
    import pwicp_b200
    ctx = pwicp_b200.Context(0)
    yield ctx
    ctx.close()


# ---- BASELINE configs[0]: the reference's shipped two-epoch pair + the result the reference recorded for it ----------
def load_refpair(patch_stats, path=None):
    """tests/golden/refpair_e2.npz (made by tests/golden/make_refpair.py) -> the centroid-level pair Piecewise_ICP holds after
    PatchGenerationAndRefinement + calBPandCTSTD (src/Registration.cpp:653-664).  patch_stats(points, offsets) supplies the
    per-patch constants: the oracle's on the CPU, the device's (pwicp_patch_stats) on the GPU."""
    import numpy as np
    z = np.load(path or os.path.join(ROOT, "tests", "golden", "refpair_e2.npz"))

    def patches(cloud, lab):
        keep = np.nonzero(lab >= 0)[0]
        order = keep[np.argsort(lab[keep], kind="stable")]            # patch k = its cloud points in cloud order
        return cloud[order], np.concatenate([[0], np.cumsum(np.bincount(lab[keep]))]).astype(np.int32)

    p1, o1 = patches(z["cloud1"], z["lab1"])
    p2, o2 = patches(z["cloud2"], z["lab2"])
    s1, s2 = patch_stats(p1, o1), patch_stats(p2, o2)
    nrm1 = s1["nrm"].copy()
    nrm1[~((np.diff(o1) > 6) & (s1["nrm_ok"] != 0))] = (0, 0, 1)     # generateCentroidCloudWithPatchNormals, src/CommonFunc.cpp:367
    res, sv, dtinit, dtmin = (float(v) for v in z["config"])
    d = {"cloud1": z["cloud1"], "ct1": s1["ct"], "nrm1": nrm1, "nrm1_ok": s1["nrm_ok"], "ctstd1": s1["ctstd"],
         "cloud2": z["cloud2"], "ct2": s2["ct"], "bp2": s2["bp"].reshape(-1, 3), "bpstd2": s2["bpstd"],
         "patch_off2": o2, "patch_pts2": p2, "Res1": res, "Res2": res, "SVRes1": sv, "SVRes2": sv, "DTmin": dtmin}
    S = np.eye(4, dtype=np.float32); S[:3, 3] = z["shift"]
    Si = np.eye(4, dtype=np.float32); Si[:3, 3] = -z["shift"]
    return {"pair": d, "DTinit": dtinit, "S": S, "Sinv": Si, "T_recorded": z["T_recorded"], "VCM_recorded": z["VCM_recorded"],
            "T_truth": z["T_truth"]}
