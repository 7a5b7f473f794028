"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class IcpParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tf_eps", C.c_double), ("fit_eps", C.c_double),
                ("force_iters", C.c_int), ("reduce_mode", C.c_int), ("group_batches", C.c_int),
                ("reserved", C.c_int), ("rot_thr_default", C.c_int)]


class Pair(C.Structure):
    _fields_ = [("cloud1", C.c_void_p), ("m1", C.c_int),
                ("ct1", C.c_void_p), ("n1", C.c_int),
                ("nrm1", C.c_void_p), ("nrm1_ok", C.c_void_p), ("ctstd1", C.c_void_p),
                ("cloud2", C.c_void_p), ("m2", C.c_int),
                ("ct2", C.c_void_p), ("n2", C.c_int),
                ("bp2", C.c_void_p), ("bpstd2", C.c_void_p),
                ("patch_off2", C.c_void_p), ("patch_pts2", C.c_void_p),
                ("Res1", C.c_float), ("Res2", C.c_float), ("SVRes1", C.c_float),
                ("SVRes2", C.c_float), ("DTmin", C.c_float)]


class State(C.Structure):
    _fields_ = [("currDT", C.c_float), ("BBchange_1", C.c_float), ("BBchange_2", C.c_float),
                ("toStage2", C.c_int), ("toStage3", C.c_int)]


class IterStats(C.Structure):
    _fields_ = [("n_stable", C.c_int), ("n_stable_pts", C.c_int), ("icp_iters", C.c_int),
                ("icp_state", C.c_int), ("LoDet_min", C.c_float), ("LoDet_max", C.c_float),
                ("maxBBchange", C.c_float), ("P75", C.c_double), ("bb6", C.c_double * 6)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pwicp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_nn.argtypes = [f32p, C.c_int, f32p, C.c_int, i32p, f32p]
        L.orc_nn_brute.argtypes = [f32p, C.c_int, f32p, C.c_int, i32p, f32p]
        L.orc_tree_build.argtypes = [f32p, C.c_int]
        L.orc_tree_build.restype = C.c_void_p
        L.orc_tree_query.argtypes = [C.c_void_p, f32p, C.c_int, i32p, f32p]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_transform.argtypes = [f32p, C.c_int, f32p]
        L.orc_lls_step.argtypes = [f32p, i32p, C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_int,
                                   f64p, f64p, f64p, f32p]
        L.orc_icp_p2plane.argtypes = [f32p, f32p, C.c_int, f32p, C.c_int, C.POINTER(IcpParams),
                                      f32p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_octree_bbox.argtypes = [f32p, C.c_int, C.c_double, f64p]
        L.orc_bbox_corner_change.argtypes = [f64p, f32p]
        L.orc_bbox_corner_change.restype = C.c_float
        L.orc_percentile_nn.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_float]
        L.orc_percentile_nn.restype = C.c_double
        L.orc_vcm.argtypes = [f32p, f32p, C.c_int, f32p, C.c_int, f64p, C.POINTER(C.c_int)]
        L.orc_patch_normal.argtypes = [f32p, C.c_int, f32p]
        L.orc_matrix2angle.argtypes = [f32p, f32p]
        L.orc_single_iteration.argtypes = [C.POINTER(Pair), C.POINTER(State), C.POINTER(IcpParams),
                                           f32p, f64p, C.POINTER(C.c_int), C.c_void_p,
                                           C.POINTER(IterStats)]
        L.orc_piecewise_icp.argtypes = [C.POINTER(Pair), C.c_int, C.c_float, C.POINTER(IcpParams),
                                        C.c_int, f32p, C.POINTER(C.c_int), f32p, f64p, C.c_void_p]
        L.orc_mat4_mul.argtypes = [f32p, f32p, f32p]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def icp_params(max_iter=100, tf_eps=1e-8, fit_eps=1e-6, force_iters=0, reduce_mode=0,
               group_batches=0, rot_thr_default=0, threads=0):
    """threads > 1: NN queries and row terms on that many threads (bench only; the reference is single-threaded)."""
    return IcpParams(max_iter, tf_eps, fit_eps, force_iters, reduce_mode, group_batches, threads, rot_thr_default)


def set_threads(n):
    """Threads for the independent NN queries of the outer iteration / percentile (results do not depend on it)."""
    lib().orc_set_threads(int(n))


def nn(tgt, qry, brute=False):
    tgt, qry = _f32(tgt), _f32(qry)
    idx = np.empty(len(qry), np.int32)
    d2 = np.empty(len(qry), np.float32)
    (lib().orc_nn_brute if brute else lib().orc_nn)(tgt, len(tgt), qry, len(qry), idx, d2)
    return idx, d2


# ---- oracle/_ref: the reference's own KD-tree (codelibrary/util/tree/kd_tree.h), compiled where it lies ----
REF_KDTREE = os.path.join(_HERE, "_ref", "libref_kdtree.so")
_REF = None


def ref_available():
    """True when oracle/_ref/libref_kdtree.so exists (built by `make -C oracle` wherever /root/reference is
    present; the built file travels to the GPU box, the reference sources do not)."""
    return os.path.exists(REF_KDTREE)


def _ref():
    global _REF
    if _REF is None:
        L = C.CDLL(REF_KDTREE)
        L.ref_kdtree_nn.argtypes = [f32p, C.c_int, f32p, C.c_int, i32p, f64p]
        L.ref_kdtree_knn.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p]
        _REF = L
    return _REF


def ref_nn(tgt, qry):
    """1-NN by the reference's KD-tree; d2 in its own (double) metric."""
    tgt, qry = _f32(tgt), _f32(qry)
    idx = np.empty(len(qry), np.int32)
    d2 = np.empty(len(qry), np.float64)
    if _ref().ref_kdtree_nn(tgt, len(tgt), qry, len(qry), idx, d2) != 0:
        raise ValueError("ref_kdtree_nn: empty target")
    return idx, d2


def ref_knn(tgt, qry, k):
    tgt, qry = _f32(tgt), _f32(qry)
    idx = np.empty((len(qry), k), np.int32)
    if _ref().ref_kdtree_knn(tgt, len(tgt), qry, len(qry), k, idx) != 0:
        raise ValueError("ref_kdtree_knn: bad k / empty target")
    return idx


def transform(pts, T):
    out = _f32(pts).copy()
    lib().orc_transform(out, len(out), _f32(T).reshape(16))
    return out


def lls_step(src, match, tgt, nrm, reduce_mode=0, group_batches=0):
    ATA = np.zeros(36); ATb = np.zeros(6); x = np.zeros(6); T = np.zeros(16, np.float32)
    lib().orc_lls_step(_f32(src), np.ascontiguousarray(match, np.int32), len(src), _f32(tgt),
                       _f32(nrm), reduce_mode, group_batches, 0, ATA, ATb, x, T)
    return ATA.reshape(6, 6), ATb, x, T.reshape(4, 4)


def icp(tgt, nrm, src, prm=None, trace=False):
    tgt, nrm, src = _f32(tgt), _f32(nrm), _f32(src)
    prm = prm or icp_params()
    T = np.zeros(16, np.float32)
    nit, cs = C.c_int(0), C.c_int(0)
    mse = np.zeros(prm.max_iter) if trace else None
    Ttr = np.zeros((prm.max_iter, 16), np.float32) if trace else None
    itr = np.zeros((prm.max_iter, len(src)), np.int32) if trace else None
    lib().orc_icp_p2plane(tgt, nrm, len(tgt), src, len(src), C.byref(prm), T, C.byref(nit),
                          C.byref(cs),
                          mse.ctypes.data if trace else None,
                          Ttr.ctypes.data if trace else None,
                          itr.ctypes.data if trace else None)
    out = {"T": T.reshape(4, 4), "n_iter": nit.value, "state": cs.value}
    if trace:
        out.update(mse=mse[:nit.value], T_trace=Ttr[:nit.value].reshape(-1, 4, 4),
                   idx_trace=itr[:nit.value])
    return out


def octree_bbox(pts, res):
    bb = np.zeros(6)
    lib().orc_octree_bbox(_f32(pts), len(pts), float(res), bb)
    return bb


def bbox_corner_change(bb6, T):
    return float(lib().orc_bbox_corner_change(np.ascontiguousarray(bb6, np.float64), _f32(T).reshape(16)))


def percentile_nn(cloud1, cloud2, pct=0.75):
    return float(lib().orc_percentile_nn(_f32(cloud1), len(cloud1), _f32(cloud2), len(cloud2), pct))


def vcm(tgt, nrm, src):
    out = np.zeros(36); s = C.c_int(0)
    lib().orc_vcm(_f32(tgt), _f32(nrm), len(tgt), _f32(src), len(src), out, C.byref(s))
    return out.reshape(6, 6), bool(s.value)


def patch_normal(pts):
    n = np.zeros(3, np.float32)
    ok = lib().orc_patch_normal(_f32(pts), len(pts), n)
    return n, bool(ok)


def patch_stats(pts, off):
    """calPatchCTandBP + calPatchNormal + calPatchSTD / calBPandCTSTD for every patch (points packed by patch)."""
    pts = _f32(pts)
    off = np.ascontiguousarray(off, np.int32)
    n = len(off) - 1
    ct = np.zeros((n, 3), np.float32); bp = np.zeros((n, 6, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
    ok = np.zeros(n, np.uint8); bs = np.zeros(n, np.float32); cs = np.zeros(n, np.float32)
    L = lib()
    L.orc_patch_stats.argtypes = [f32p, i32p, C.c_int, f32p, f32p, f32p, np.ctypeslib.ndpointer(np.uint8, flags="C"), f32p, f32p]
    L.orc_patch_stats.restype = None
    L.orc_patch_stats(pts, off, n, ct, bp, nrm, ok, bs, cs)
    return {"ct": ct, "bp": bp, "nrm": nrm, "nrm_ok": ok, "bpstd": bs, "ctstd": cs}


def voxel_grid(xyz, leaf, msvc_order=False):
    """pcl::VoxelGrid with a cubic leaf (F4).  msvc_order: points of a voxel summed in the order the Microsoft STL's
    std::sort leaves them (the reference's Windows build) instead of the input order."""
    p = _f32(xyz)
    out = np.zeros_like(p)
    L = lib()
    fn = L.orc_voxel_grid_msvc if msvc_order else L.orc_voxel_grid
    fn.argtypes = [f32p, C.c_int, C.c_float, f32p]
    m = fn(p, len(p), leaf, out)
    return out[:m].copy()


def knn_mean_dist(xyz, k):
    """first pass of pcl::StatisticalOutlierRemoval: mean distance to the k nearest other points (F4)."""
    p = _f32(xyz)
    out = np.zeros(len(p), np.float32)
    L = lib()
    L.orc_knn_mean_dist.argtypes = [f32p, C.c_int, C.c_int, f32p]
    if L.orc_knn_mean_dist(p, len(p), k, out) != 0:
        raise ValueError("knn_mean_dist: need 1 <= k < n")
    return out


def sor_select(xyz, mean_dist, std_mult):
    p = _f32(xyz)
    md = _f32(mean_dist)
    out = np.zeros_like(p)
    thr = C.c_double(0)
    L = lib()
    L.orc_sor_select.argtypes = [f32p, C.c_int, f32p, C.c_double, f32p, C.POINTER(C.c_double)]
    m = L.orc_sor_select(p, len(p), md, std_mult, out, C.byref(thr))
    return out[:m].copy(), thr.value


def preprocess(xyz, leaf, k=14, std_mult=5.0, msvc_order=False):
    """PCpreprocessing(cloud, out, true, leaf, k, std_mult) (src/CommonFunc.cpp:423-439)."""
    v = voxel_grid(xyz, leaf, msvc_order)
    return sor_select(v, knn_mean_dist(v, k), std_mult)[0]


def knn_normals(xyz, k=45):
    """k nearest neighbours (self first) + PCA normal of every point, as the reference's segmentation front end computes them."""
    p = _f32(xyz)
    nb = np.zeros((len(p), k), np.int32)
    nrm = np.zeros((len(p), 3), np.float64)
    L = lib()
    L.orc_knn_normals.argtypes = [f32p, C.c_int, C.c_int, i32p, f64p]
    if L.orc_knn_normals(p, len(p), k, nb.reshape(-1), nrm.reshape(-1)) != 0:
        raise ValueError("knn_normals: need 1 <= k <= n")
    return nb, nrm


def ref_knn_normals(xyz, k=45):
    """the same through the reference's own code (oracle/_ref/libref_supervoxel.so)"""
    p = _f32(xyz)
    nb = np.zeros((len(p), k), np.int32)
    nrm = np.zeros((len(p), 3), np.float64)
    L = C.CDLL(os.path.join(_HERE, "_ref", "libref_supervoxel.so"))
    L.ref_knn_normals.argtypes = [f32p, C.c_int, C.c_int, i32p, f64p]
    if L.ref_knn_normals(p, len(p), k, nb.reshape(-1), nrm.reshape(-1)) != 0:
        raise ValueError("ref_knn_normals: need n > k")
    return nb, nrm


def matrix2angle(T):
    a = np.zeros(3, np.float32)
    lib().orc_matrix2angle(_f32(T).reshape(16), a)
    return a


def mat4_mul(A, B):
    out = np.zeros(16, np.float32)
    lib().orc_mat4_mul(_f32(A).reshape(16), _f32(B).reshape(16), out)
    return out.reshape(4, 4)


class PairData:
    """Owns numpy copies of a centroid-level pair and exposes them as an orc_pair."""

    def __init__(self, d):
        g = lambda k, dt=np.float32: np.ascontiguousarray(d[k], dtype=dt).copy()
        self.cloud1, self.ct1, self.nrm1, self.ctstd1 = g("cloud1"), g("ct1"), g("nrm1"), g("ctstd1")
        self.cloud2, self.ct2, self.bp2, self.bpstd2 = g("cloud2"), g("ct2"), g("bp2"), g("bpstd2")
        self.patch_off2 = g("patch_off2", np.int32)
        self.patch_pts2 = g("patch_pts2")
        self.nrm1_ok = g("nrm1_ok", np.uint8) if d.get("nrm1_ok") is not None else None
        p = Pair()
        p.cloud1, p.m1 = self.cloud1.ctypes.data, len(self.cloud1)
        p.ct1, p.n1 = self.ct1.ctypes.data, len(self.ct1)
        p.nrm1, p.ctstd1 = self.nrm1.ctypes.data, self.ctstd1.ctypes.data
        p.nrm1_ok = self.nrm1_ok.ctypes.data if self.nrm1_ok is not None else None
        p.cloud2, p.m2 = self.cloud2.ctypes.data, len(self.cloud2)
        p.ct2, p.n2 = self.ct2.ctypes.data, len(self.ct2)
        p.bp2, p.bpstd2 = self.bp2.ctypes.data, self.bpstd2.ctypes.data
        p.patch_off2, p.patch_pts2 = self.patch_off2.ctypes.data, self.patch_pts2.ctypes.data
        p.Res1, p.Res2 = d["Res1"], d["Res2"]
        p.SVRes1, p.SVRes2, p.DTmin = d["SVRes1"], d["SVRes2"], d["DTmin"]
        self.c = p


def single_iteration(pd, state, prm=None):
    T = np.zeros(16, np.float32); V = np.zeros(36); vw = C.c_int(0)
    flags = np.zeros(pd.c.n2, np.uint8); stats = IterStats()
    rc = lib().orc_single_iteration(C.byref(pd.c), C.byref(state), C.byref(prm) if prm else None,
                                    T, V, C.byref(vw), flags.ctypes.data, C.byref(stats))
    return rc, T.reshape(4, 4), (V.reshape(6, 6) if vw.value else None), flags, stats


def piecewise_icp(pd, manual_dt, DTinit, prm=None, max_outer=200):
    series = np.zeros(max_outer + 1, np.float32); ns = C.c_int(0)
    T = np.zeros(16, np.float32); V = np.zeros(36)
    stats = (IterStats * max_outer)()
    rc = lib().orc_piecewise_icp(C.byref(pd.c), int(manual_dt), DTinit, C.byref(prm) if prm else None,
                                 max_outer, series, C.byref(ns), T, V, C.cast(stats, C.c_void_p))
    return {"rc": rc, "DTseries": series[:ns.value].copy(), "T": T.reshape(4, 4),
            "VCM": V.reshape(6, 6), "stats": list(stats[:max(rc, 0)])}
