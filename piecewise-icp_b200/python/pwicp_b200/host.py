"""ctypes binding of libpwicp_host.so: the reference's two public entry points (the same calls
python/main.py:21-41 makes against the Windows DLL) and the epoch-sharded 4D form."""
import ctypes as C
import os

import numpy as np

from . import ROOT

LIB_PATH = os.path.join(ROOT, "libpwicp_host.so")
_lib = None


class EpochRecord(C.Structure):
    _fields_ = [("step", C.c_int), ("status", C.c_int), ("time_stamp", C.c_longlong),
                ("T", C.c_float * 16), ("para", C.c_float * 6), ("VCM", C.c_double * 36),
                ("seconds", C.c_float), ("pad", C.c_int)]


class ConfigC(C.Structure):
    _fields_ = [("path1", C.c_char * 1024), ("path2", C.c_char * 1024), ("isSetResSVsize", C.c_int),
                ("PCres1", C.c_float), ("PCres2", C.c_float), ("SVsize1", C.c_float), ("SVsize2", C.c_float),
                ("isSetDTinit", C.c_int), ("DTinit", C.c_float), ("DTmin", C.c_float), ("isVisual", C.c_int)]


assert C.sizeof(EpochRecord) == 400


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make -C {ROOT}`")
        L = C.CDLL(LIB_PATH)
        L.PiecewiseICP_pair_call.argtypes = [C.c_char_p, C.c_char_p]
        L.PiecewiseICP_pair_call.restype = C.c_bool
        L.PiecewiseICP_4D_call.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float]
        L.PiecewiseICP_4D_call.restype = C.c_bool
        L.PiecewiseICP_4D_shard.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                            C.c_int, C.POINTER(EpochRecord)]
        L.PiecewiseICP_4D_finalize.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(EpochRecord)]
        L.PiecewiseICP_4D_finalize.restype = C.c_bool
        L.pwicp_host_set_device.argtypes = [C.c_int]
        L.pwicp_host_read_config.argtypes = [C.c_char_p, C.POINTER(ConfigC)]
        L.pwicp_host_patch_normal.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pwicp_host_load_pcd.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.pwicp_host_save_pcd.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.pwicp_host_list_epochs.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.pwicp_host_patches.argtypes = [C.c_void_p, C.c_int, C.c_float] + [C.c_void_p] * 4 + [C.c_int]
        L.pwicp_host_chain_to_reference.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        L.pwicp_host_chain_to_reference.restype = None
        L.pwicp_host_register_clouds.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                                 C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def pair_call(confile, outprefix):
    return bool(lib().PiecewiseICP_pair_call(confile.encode(), outprefix.encode()))


def call_4d(confile, start_epoch, epoch_num, pair_mode, overlap_thd=0.75):
    return bool(lib().PiecewiseICP_4D_call(confile.encode(), start_epoch, epoch_num, pair_mode, overlap_thd))


def shard_4d(confile, start_epoch, epoch_num, pair_mode, overlap_thd, rank, world, device=-1):
    recs = (EpochRecord * epoch_num)()
    done = lib().PiecewiseICP_4D_shard(confile.encode(), start_epoch, epoch_num, pair_mode, overlap_thd, rank, world,
                                       device, recs)
    return done, recs


def finalize_4d(confile, start_epoch, epoch_num, pair_mode, recs):
    return bool(lib().PiecewiseICP_4D_finalize(confile.encode(), start_epoch, epoch_num, pair_mode, recs))


def records_to_array(recs):
    return np.frombuffer(bytes(recs), dtype=np.uint8).reshape(len(recs), C.sizeof(EpochRecord)).copy()


def array_to_records(arr):
    n = arr.shape[0]
    recs = (EpochRecord * n)()
    C.memmove(recs, np.ascontiguousarray(arr).ctypes.data, n * C.sizeof(EpochRecord))
    return recs


def merge_records(per_rank_arrays):
    """Merge the record arrays of all ranks: entry k is taken from the rank that registered it."""
    out = per_rank_arrays[0].copy()
    n = out.shape[0]
    for arr in per_rank_arrays[1:]:
        for k in range(n):
            status = int(np.frombuffer(arr[k, 4:8].tobytes(), np.int32)[0])
            if status == 1:
                out[k] = arr[k]
    return out


def gather_records(recs, dist=None, device=None):
    """All-gather of the fixed-size per-epoch records (the only collective of the 4D mode,
    SURVEY.md 8e): NCCL when `device` is a CUDA device, gloo on CPU."""
    arr = records_to_array(recs)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return arr
    import torch
    t = torch.from_numpy(arr)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return merge_records([o.cpu().numpy() for o in outs])


def read_config(path):
    c = ConfigC()
    ok = lib().pwicp_host_read_config(path.encode(), C.byref(c))
    if not ok:
        return None
    return {"FolderFilePath1": c.path1.decode(), "FolderFilePath2": c.path2.decode(),
            "isSetResSVsize": bool(c.isSetResSVsize), "PCres1": c.PCres1, "PCres2": c.PCres2,
            "SVsize1": c.SVsize1, "SVsize2": c.SVsize2, "isSetDTinit": bool(c.isSetDTinit),
            "DTinit": c.DTinit, "DTmin": c.DTmin, "isVisual": bool(c.isVisual)}


def patch_normal(pts):
    p = np.ascontiguousarray(pts, np.float32)
    n = np.zeros(3, np.float32)
    ok = lib().pwicp_host_patch_normal(p.ctypes.data, len(p), n.ctypes.data)
    return n, bool(ok)


def load_pcd(path):
    n = lib().pwicp_host_load_pcd(path.encode(), None, 0)
    if n < 0:
        raise IOError(path)
    xyz = np.zeros((n, 3), np.float32)
    lib().pwicp_host_load_pcd(path.encode(), xyz.ctypes.data, n)
    return xyz


def save_pcd(path, xyz):
    p = np.ascontiguousarray(xyz, np.float32)
    return lib().pwicp_host_save_pcd(path.encode(), p.ctypes.data, len(p)) == 0


def list_epochs(folder):
    t = np.zeros(4096, np.int64)
    n = lib().pwicp_host_list_epochs(folder.encode(), t.ctypes.data, len(t))
    return t[:n].tolist()


def patches(xyz, sv_res, cap=1 << 20):
    p = np.ascontiguousarray(xyz, np.float32)
    ct = np.zeros((cap, 3), np.float32); bp = np.zeros((cap, 18), np.float32)
    sb = np.zeros(cap, np.float32); sc = np.zeros(cap, np.float32)
    n = lib().pwicp_host_patches(p.ctypes.data, len(p), sv_res, ct.ctypes.data, bp.ctypes.data, sb.ctypes.data, sc.ctypes.data, cap)
    return {"ct": ct[:n], "bp": bp[:n].reshape(-1, 3), "bpstd": sb[:n], "ctstd": sc[:n]}


def preprocess(xyz, leaf, k=14, mult=5.0, downsample=True, device=True):
    """PCpreprocessing (src/CommonFunc.cpp:423-439): on the device like the drivers (device=True) or the host statements."""
    p = np.ascontiguousarray(xyz, np.float32)
    out = np.zeros_like(p)
    L = lib()
    L.pwicp_host_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_int]
    m = L.pwicp_host_preprocess(p.ctypes.data, len(p), int(downsample), leaf, k, mult, int(device), out.ctypes.data, len(p))
    return out[:m].copy()


def register_clouds(xyz1, xyz2, res, sv, dtinit, dtmin, mode=0):
    """Piecewise_ICP on two in-memory clouds (already pre-processed and shifted).  mode 0: the
    device outer loop; mode 1: the reference's while(!stage3) PwICP_singleIteration loop."""
    a, b = np.ascontiguousarray(xyz1, np.float32), np.ascontiguousarray(xyz2, np.float32)
    T = np.zeros(16, np.float32); V = np.zeros(36); s = np.zeros(256, np.float32)
    n = lib().pwicp_host_register_clouds(a.ctypes.data, len(a), b.ctypes.data, len(b), res, sv, dtinit, dtmin, mode,
                                         T.ctypes.data, V.ctypes.data, s.ctypes.data, len(s))
    return {"T": T.reshape(4, 4), "VCM": V.reshape(6, 6), "DTseries": s[:n].copy()}
