"""pwicp_b200 -- thin ctypes binding of libpwicp.so (the C ABI in include/pwicp.h).

The library is the product; this module only marshals numpy arrays into the C ABI, exactly the way
the reference's python/main.py:10-41 binds its DLL with ctypes.  There is no CPU fallback: if the
CUDA extension is missing or no CUDA device is visible, every compute call raises.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(_PKG, "..", ".."))          # piecewise-icp_b200/
LIB_PATH = os.path.join(ROOT, "libpwicp.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

CONV_NAMES = {0: "none", 1: "iterations", 2: "transform", 3: "abs_mse", 4: "rel_mse", 5: "no_corr"}
TGT_CENTROIDS, TGT_CLOUD1 = 0, 1


class PwicpError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"pwicp status {status}: {msg}")
        self.status = status


class IcpParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tf_eps", C.c_double), ("fit_eps", C.c_double),
                ("force_iters", C.c_int), ("rot_thr_default", C.c_int)]


class IcpResult(C.Structure):
    _fields_ = [("n_iter", C.c_int), ("conv_state", C.c_int), ("grid_blocks", C.c_int),
                ("warps_per_block", C.c_int), ("group_batches", C.c_int), ("device_ms", C.c_float),
                ("correspondences", C.c_longlong), ("kernel_ms", C.c_float), ("natural_iters", C.c_int), ("natural_state", C.c_int),
                ("sort_ms", C.c_float), ("prepass_ms", C.c_float), ("research_ms", C.c_float)]


class PairParams(C.Structure):
    _fields_ = [("Res1", C.c_float), ("Res2", C.c_float), ("SVRes1", C.c_float),
                ("SVRes2", C.c_float), ("DTmin", C.c_float)]


class State(C.Structure):
    _fields_ = [("currDT", C.c_float), ("BBchange_1", C.c_float), ("BBchange_2", C.c_float),
                ("toStage2", C.c_int), ("toStage3", C.c_int)]


class IterStats(C.Structure):
    _fields_ = [("n_stable", C.c_int), ("n_stable_pts", C.c_int), ("icp_iters", C.c_int),
                ("icp_state", C.c_int), ("LoDet_min", C.c_float), ("LoDet_max", C.c_float),
                ("maxBBchange", C.c_float), ("P75", C.c_double), ("bb6", C.c_double * 6),
                ("vcm_written", C.c_int), ("vcm_singular", C.c_int), ("device_ms", C.c_float)]


EXPORTS = [
    "pwicp_version", "pwicp_device_count", "pwicp_ctx_create", "pwicp_ctx_destroy",
    "pwicp_last_error", "pwicp_last_device_ms", "pwicp_launch_count", "pwicp_flush_l2",
    "pwicp_sync", "pwicp_set_cells_per_point", "pwicp_target_upload", "pwicp_target_rebuild", "pwicp_source_upload",
    "pwicp_clouds_upload", "pwicp_source_download", "pwicp_nn", "pwicp_icp_default_params",
    "pwicp_icp_source_upload", "pwicp_icp_source_all", "pwicp_icp_run", "pwicp_icp_order",
    "pwicp_icp_p2plane",
    "pwicp_single_iteration", "pwicp_piecewise_icp", "pwicp_percentile_nn", "pwicp_overlap_ratio",
    "pwicp_self_nn", "pwicp_vcm", "pwicp_transform", "pwicp_octree_bbox", "pwicp_bbox_corner_change",
    "pwicp_matrix2angle", "pwicp_mat4_mul", "pwicp_patch_stats",
    "pwicp_voxel_grid", "pwicp_knn_mean_dist", "pwicp_preprocess", "pwicp_last_knn_kernel_ms", "pwicp_knn_normals",
    "pwicp_icp_profile", "pwicp_icp_phase_profile",
]

_lib = None


def load_library(path=None):
    """Load libpwicp.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `make -C {ROOT}` "
                          "(pwicp_b200 has no CPU fallback)")
    L = C.CDLL(p)
    vp = C.c_void_p
    L.pwicp_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.pwicp_ctx_destroy.argtypes = [vp]
    L.pwicp_ctx_destroy.restype = None
    L.pwicp_last_error.argtypes = [vp]
    L.pwicp_last_error.restype = C.c_char_p
    L.pwicp_last_device_ms.argtypes = [vp]
    L.pwicp_last_device_ms.restype = C.c_float
    L.pwicp_launch_count.argtypes = [vp]
    L.pwicp_launch_count.restype = C.c_longlong
    L.pwicp_flush_l2.argtypes = [vp]
    L.pwicp_sync.argtypes = [vp]
    L.pwicp_set_cells_per_point.argtypes = [vp, C.c_float]
    L.pwicp_target_upload.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    L.pwicp_target_rebuild.argtypes = [vp]
    L.pwicp_source_upload.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
    L.pwicp_clouds_upload.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.pwicp_source_download.argtypes = [vp, vp, vp, vp, vp]
    L.pwicp_nn.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp]
    L.pwicp_icp_default_params.argtypes = [C.POINTER(IcpParams)]
    L.pwicp_icp_default_params.restype = None
    L.pwicp_icp_source_upload.argtypes = [vp, vp, C.c_int]
    L.pwicp_icp_source_all.argtypes = [vp]
    L.pwicp_icp_run.argtypes = [vp, C.POINTER(IcpParams), vp, C.POINTER(IcpResult), vp, vp, vp]
    L.pwicp_icp_order.argtypes = [vp, vp]
    L.pwicp_icp_profile.argtypes = [vp, vp, vp, C.c_int]
    L.pwicp_icp_phase_profile.argtypes = [vp, vp, C.c_int]
    L.pwicp_icp_p2plane.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.POINTER(IcpParams), vp,
                                    C.POINTER(IcpResult)]
    L.pwicp_single_iteration.argtypes = [vp, C.POINTER(PairParams), C.POINTER(State),
                                         C.POINTER(IcpParams), vp, vp, vp, C.POINTER(IterStats)]
    L.pwicp_piecewise_icp.argtypes = [vp, C.POINTER(PairParams), C.c_int, C.c_float,
                                      C.POINTER(IcpParams), C.c_int, vp, C.POINTER(C.c_int), vp, vp,
                                      C.POINTER(C.c_int), vp]
    L.pwicp_percentile_nn.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_float, C.POINTER(C.c_double)]
    L.pwicp_overlap_ratio.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_float, C.POINTER(C.c_float)]
    L.pwicp_self_nn.argtypes = [vp, vp, C.c_int, vp]
    L.pwicp_patch_stats.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.pwicp_knn_normals.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.pwicp_last_knn_kernel_ms.argtypes = [vp]
    L.pwicp_last_knn_kernel_ms.restype = C.c_float
    L.pwicp_voxel_grid.argtypes = [vp, vp, C.c_int, C.c_float, vp, C.POINTER(C.c_int)]
    L.pwicp_knn_mean_dist.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.pwicp_preprocess.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, vp, C.POINTER(C.c_int)]
    L.pwicp_vcm.argtypes = [vp, vp, C.c_int, vp, C.POINTER(C.c_int)]
    L.pwicp_transform.argtypes = [vp, vp, C.c_int, vp]
    L.pwicp_octree_bbox.argtypes = [vp, vp, C.c_int, C.c_double, vp]
    L.pwicp_bbox_corner_change.argtypes = [_f64p, _f32p]
    L.pwicp_bbox_corner_change.restype = C.c_float
    L.pwicp_matrix2angle.argtypes = [_f32p, _f32p]
    L.pwicp_matrix2angle.restype = None
    L.pwicp_mat4_mul.argtypes = [_f32p, _f32p, _f32p]
    L.pwicp_mat4_mul.restype = None
    if path is None:
        _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data


def icp_params(max_iter=100, tf_eps=1e-8, fit_eps=1e-6, force_iters=0, rot_thr_default=0):
    return IcpParams(max_iter, tf_eps, fit_eps, force_iters, rot_thr_default)


def matrix2angle(T):
    a = np.zeros(3, np.float32)
    load_library().pwicp_matrix2angle(_f32(T).reshape(16), a)
    return a


def bbox_corner_change(bb6, T):
    return float(load_library().pwicp_bbox_corner_change(np.ascontiguousarray(bb6, np.float64),
                                                         _f32(T).reshape(16)))


def mat4_mul(A, B):
    out = np.zeros(16, np.float32)
    load_library().pwicp_mat4_mul(_f32(A).reshape(16), _f32(B).reshape(16), out)
    return out.reshape(4, 4)


class Context:
    """One device context (one CUDA stream, all device buffers of a pair)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        st = self.L.pwicp_ctx_create(device, C.byref(h))
        if st != 0:
            raise PwicpError(st, self.L.pwicp_last_error(None).decode())
        self.h = h
        self.n1 = self.n2 = self.m1 = self.m2 = self.mp2 = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.pwicp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, st):
        if st != 0:
            raise PwicpError(st, self.L.pwicp_last_error(self.h).decode())

    # -- plumbing
    def last_device_ms(self):
        return float(self.L.pwicp_last_device_ms(self.h))

    def launch_count(self):
        return int(self.L.pwicp_launch_count(self.h))

    def flush_l2(self):
        self._chk(self.L.pwicp_flush_l2(self.h))

    def sync(self):
        self._chk(self.L.pwicp_sync(self.h))

    def set_cells_per_point(self, cpp):
        self._chk(self.L.pwicp_set_cells_per_point(self.h, cpp))

    # -- uploads
    def target_upload(self, ct, nrm=None, ctstd=None, nrm_ok=None):
        ct = _f32(ct)
        nrm = None if nrm is None else _f32(nrm)
        ctstd = None if ctstd is None else _f32(ctstd)
        ok = None if nrm_ok is None else np.ascontiguousarray(nrm_ok, np.uint8)
        self._chk(self.L.pwicp_target_upload(self.h, _ptr(ct), _ptr(nrm), _ptr(ok), _ptr(ctstd), len(ct)))
        self.n1 = len(ct)

    def target_rebuild(self):
        self._chk(self.L.pwicp_target_rebuild(self.h))
        return self.last_device_ms()

    def source_upload(self, ct, bp=None, bpstd=None, patch_off=None, patch_xyz=None):
        ct = _f32(ct)
        bp = None if bp is None else _f32(bp)
        bpstd = None if bpstd is None else _f32(bpstd)
        off = None if patch_off is None else np.ascontiguousarray(patch_off, np.int32)
        pp = None if patch_xyz is None else _f32(patch_xyz)
        self._chk(self.L.pwicp_source_upload(self.h, _ptr(ct), _ptr(bp), _ptr(bpstd), _ptr(off), _ptr(pp), len(ct)))
        self.n2 = len(ct)
        self.mp2 = 0 if off is None else int(off[-1])

    def clouds_upload(self, cloud1, cloud2):
        """cloud1 = None keeps the resident cloud1 and its grid (the reference epoch of a series)."""
        c2 = _f32(cloud2)
        if cloud1 is None:
            self._chk(self.L.pwicp_clouds_upload(self.h, None, 0, _ptr(c2), len(c2)))
        else:
            c1 = _f32(cloud1)
            self._chk(self.L.pwicp_clouds_upload(self.h, _ptr(c1), len(c1), _ptr(c2), len(c2)))
            self.m1 = len(c1)
        self.m2 = len(c2)

    def upload_pair(self, d):
        self.target_upload(d["ct1"], d["nrm1"], d["ctstd1"], d.get("nrm1_ok"))
        self.source_upload(d["ct2"], d["bp2"], d["bpstd2"], d["patch_off2"], d["patch_pts2"])
        self.clouds_upload(d["cloud1"], d["cloud2"])

    def upload_source_side(self, d):
        """The moving epoch of a pair whose reference side (pwicp_target_upload, cloud1) is already resident."""
        self.source_upload(d["ct2"], d["bp2"], d["bpstd2"], d["patch_off2"], d["patch_pts2"])
        self.clouds_upload(None, d["cloud2"])

    def source_download(self):
        cloud2 = np.zeros((self.m2, 3), np.float32)
        ct = np.zeros((self.n2, 3), np.float32)
        bp = np.zeros((6 * self.n2, 3), np.float32)
        pp = np.zeros((self.mp2, 3), np.float32)
        self._chk(self.L.pwicp_source_download(self.h, _ptr(cloud2), _ptr(ct), _ptr(bp), _ptr(pp)))
        return {"cloud2": cloud2, "ct2": ct, "bp2": bp, "patch_pts2": pp}

    # -- A1
    def nn(self, qry, which=TGT_CENTROIDS):
        q = _f32(qry)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        self._chk(self.L.pwicp_nn(self.h, which, _ptr(q), len(q), _ptr(idx), _ptr(d2)))
        return idx, d2

    # -- A3-A6
    def icp_source_upload(self, src):
        s = _f32(src)
        self._chk(self.L.pwicp_icp_source_upload(self.h, _ptr(s), len(s)))
        self._n_icp = len(s)

    def icp_source_all(self):
        self._chk(self.L.pwicp_icp_source_all(self.h))
        self._n_icp = self.n2

    def icp_run(self, prm=None, trace=False):
        prm = prm or icp_params()
        T = np.zeros(16, np.float32)
        res = IcpResult()
        mse = Ttr = itr = None
        if trace:
            mse = np.zeros(prm.max_iter)
            Ttr = np.zeros((prm.max_iter, 16), np.float32)
            itr = np.zeros((prm.max_iter, self._n_icp), np.int32)
        self._chk(self.L.pwicp_icp_run(self.h, C.byref(prm), _ptr(T), C.byref(res), _ptr(mse), _ptr(Ttr), _ptr(itr)))
        out = {"T": T.reshape(4, 4), "n_iter": res.n_iter, "state": res.conv_state,
               "grid_blocks": res.grid_blocks, "warps_per_block": res.warps_per_block,
               "group_batches": res.group_batches, "natural_iters": res.natural_iters,
               "natural_state": res.natural_state, "sort_ms": res.sort_ms, "prepass_ms": res.prepass_ms, "research_ms": res.research_ms,
               "device_ms": res.device_ms, "kernel_ms": res.kernel_ms, "correspondences": res.correspondences}
        if trace:
            out.update(mse=mse[:res.n_iter], T_trace=Ttr[:res.n_iter].reshape(-1, 4, 4),
                       idx_trace=itr[:res.n_iter])
        return out

    def icp_profile(self, cap=1024):
        """Per-iteration profile of the last inner loop: (microseconds per iteration, queries that ran the search)."""
        us = np.zeros(cap)
        srch = np.zeros(cap, np.int32)
        m = self.L.pwicp_icp_profile(self.h, _ptr(us), _ptr(srch), cap)
        return us[:m], srch[:m]

    def icp_phase_profile(self, cap=1024):
        """Where CTA 0 spent each iteration of the last inner loop: (iterations, 4) microseconds since the iteration began."""
        us = np.zeros((cap, 4))
        m = self.L.pwicp_icp_phase_profile(self.h, _ptr(us), cap)
        return us[:m]

    def icp_order(self):
        perm = np.zeros(self._n_icp, np.int32)
        self._chk(self.L.pwicp_icp_order(self.h, _ptr(perm)))
        return perm

    def icp_p2plane(self, tgt, nrm, src, prm=None):
        """Host buffers in, transformation out: P2PICPwithPatchNormal(target, source, eps)."""
        prm = prm or icp_params()
        t, n, s = _f32(tgt), _f32(nrm), _f32(src)
        T = np.zeros(16, np.float32)
        res = IcpResult()
        self._chk(self.L.pwicp_icp_p2plane(self.h, _ptr(t), _ptr(n), len(t), _ptr(s), len(s), C.byref(prm), _ptr(T), C.byref(res)))
        self.n1 = len(t)
        self._n_icp = len(s)
        return {"T": T.reshape(4, 4), "n_iter": res.n_iter, "state": res.conv_state,
                "device_ms": res.device_ms, "kernel_ms": res.kernel_ms, "correspondences": res.correspondences,
                "grid_blocks": res.grid_blocks, "warps_per_block": res.warps_per_block,
                "group_batches": res.group_batches, "natural_iters": res.natural_iters,
                "natural_state": res.natural_state, "sort_ms": res.sort_ms, "prepass_ms": res.prepass_ms, "research_ms": res.research_ms}

    # -- outer iteration / loop
    def single_iteration(self, pp, state, prm=None, want_flags=True):
        T = np.zeros(16, np.float32)
        V = np.zeros(36)
        flags = np.zeros(self.n2, np.uint8) if want_flags else None
        stats = IterStats()
        self._chk(self.L.pwicp_single_iteration(self.h, C.byref(pp), C.byref(state), C.byref(prm) if prm else None,
                                                _ptr(T), _ptr(V), _ptr(flags), C.byref(stats)))
        return T.reshape(4, 4), (V.reshape(6, 6) if stats.vcm_written else None), flags, stats

    def piecewise_icp(self, pp, manual_dt, DTinit, prm=None, max_outer=200):
        series = np.zeros(max_outer + 1, np.float32)
        ns, no = C.c_int(0), C.c_int(0)
        T = np.zeros(16, np.float32)
        V = np.zeros(36)
        stats = (IterStats * max_outer)()
        st = self.L.pwicp_piecewise_icp(self.h, C.byref(pp), int(manual_dt), DTinit, C.byref(prm) if prm else None,
                                        max_outer, _ptr(series), C.byref(ns), _ptr(T), _ptr(V), C.byref(no),
                                        C.cast(stats, C.c_void_p))
        self._chk(st)
        return {"n_outer": no.value, "DTseries": series[:ns.value].copy(), "T": T.reshape(4, 4),
                "VCM": V.reshape(6, 6), "stats": list(stats[:no.value]), "device_ms": self.last_device_ms()}

    # -- stand-alone pieces
    def percentile_nn(self, cloud1, cloud2, pct=0.75):
        c1, c2 = _f32(cloud1), _f32(cloud2)
        out = C.c_double(0)
        self._chk(self.L.pwicp_percentile_nn(self.h, _ptr(c1), len(c1), _ptr(c2), len(c2), pct, C.byref(out)))
        return out.value

    def overlap_ratio(self, cloud1, cloud2, DTinit):
        c1, c2 = _f32(cloud1), _f32(cloud2)
        out = C.c_float(0)
        self._chk(self.L.pwicp_overlap_ratio(self.h, _ptr(c1), len(c1), _ptr(c2), len(c2), DTinit, C.byref(out)))
        return out.value

    def knn_normals(self, pts, k=45):
        """Indices of the k nearest points of every point (itself first) and the PCA normal over them (double)."""
        p = _f32(pts)
        nb = np.zeros((len(p), k), np.int32)
        nr = np.zeros((len(p), 3), np.float64)
        self._chk(self.L.pwicp_knn_normals(self.h, _ptr(p), len(p), k, _ptr(nb), _ptr(nr)))
        return nb, nr

    def self_nn(self, pts):
        p = _f32(pts)
        d2 = np.zeros(len(p), np.float32)
        self._chk(self.L.pwicp_self_nn(self.h, _ptr(p), len(p), _ptr(d2)))
        return d2

    def patch_stats(self, patch_xyz, patch_off):
        """Constants of every planar patch in one launch: calPatchCTandBP + calPatchNormal + calPatchSTD.
        Returns dict(ct (n,3), bp (n,6,3), nrm (n,3), nrm_ok (n,), bpstd (n,), ctstd (n,))."""
        xyz = _f32(patch_xyz)
        off = np.ascontiguousarray(patch_off, np.int32)
        n = len(off) - 1
        ct = np.zeros((n, 3), np.float32); bp = np.zeros((n, 6, 3), np.float32); nrm = np.zeros((n, 3), np.float32)
        ok = np.zeros(n, np.uint8); bs = np.zeros(n, np.float32); cs = np.zeros(n, np.float32)
        self._chk(self.L.pwicp_patch_stats(self.h, _ptr(xyz), _ptr(off), n, _ptr(ct), _ptr(bp), _ptr(nrm), _ptr(ok),
                                           _ptr(bs), _ptr(cs)))
        return {"ct": ct, "bp": bp, "nrm": nrm, "nrm_ok": ok, "bpstd": bs, "ctstd": cs}

    # -- F4: PCpreprocessing
    def voxel_grid(self, xyz, leaf):
        p = _f32(xyz)
        out = np.zeros_like(p)
        m = C.c_int(0)
        self._chk(self.L.pwicp_voxel_grid(self.h, _ptr(p), len(p), C.c_float(leaf), _ptr(out), C.byref(m)))
        return out[:m.value].copy()

    def last_knn_kernel_ms(self):
        return float(self.L.pwicp_last_knn_kernel_ms(self.h))

    def knn_mean_dist(self, xyz, k):
        p = _f32(xyz)
        out = np.zeros(len(p), np.float32)
        self._chk(self.L.pwicp_knn_mean_dist(self.h, _ptr(p), len(p), int(k), _ptr(out)))
        return out

    def preprocess(self, xyz, leaf, k=14, std_mult=5.0, downsample=True):
        p = _f32(xyz)
        out = np.zeros_like(p)
        m = C.c_int(0)
        self._chk(self.L.pwicp_preprocess(self.h, _ptr(p), len(p), int(bool(downsample)), C.c_float(leaf), int(k),
                                          C.c_double(std_mult), _ptr(out), C.byref(m)))
        return out[:m.value].copy()

    def vcm(self, src):
        s = _f32(src)
        V = np.zeros(36)
        sing = C.c_int(0)
        self._chk(self.L.pwicp_vcm(self.h, _ptr(s), len(s), _ptr(V), C.byref(sing)))
        return V.reshape(6, 6), bool(sing.value)

    def transform(self, pts, T):
        p = _f32(pts).copy()
        t = _f32(T).reshape(16)
        self._chk(self.L.pwicp_transform(self.h, _ptr(p), len(p), _ptr(t)))
        return p

    def octree_bbox(self, pts, res):
        p = _f32(pts)
        bb = np.zeros(6)
        self._chk(self.L.pwicp_octree_bbox(self.h, _ptr(p), len(p), float(res), _ptr(bb)))
        return bb
