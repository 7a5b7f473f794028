"""CPU tests of the host-side mirror (libpwicp_host.so): config parsing, file listing, PCD I/O,
patch normals, chaining to the reference epoch, and the epoch-record gather over gloo (world 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from pwicp_b200 import host, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_exports_declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pwicp_host.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    syms = sorted(set(re.findall(r"\b((?:pwicp_host_|PiecewiseICP_)\w+)\s*\(", txt)))
    assert {"PiecewiseICP_pair_call", "PiecewiseICP_4D_call", "PiecewiseICP_4D_shard", "PiecewiseICP_4D_finalize"} <= set(syms)
    L = host.lib()
    for s in syms:
        assert hasattr(L, s), s


def test_read_config_crlf_and_range_checks(tmp_path):
    p = str(tmp_path / "cfg.txt")
    synth.write_config(p, "data/in folder/", "out/", res=0.005, sv=0.05, dtinit=0.05, dtmin=0.004, crlf=True)
    c = host.read_config(p)
    assert c["FolderFilePath1"] == "data/in folder/" and c["FolderFilePath2"] == "out/"      # no stray \r
    assert c["isSetResSVsize"] and abs(c["PCres1"] - 0.005) < 1e-9 and abs(c["SVsize2"] - 0.05) < 1e-9
    assert c["isSetDTinit"] and abs(c["DTinit"] - 0.05) < 1e-9 and abs(c["DTmin"] - 0.004) < 1e-9 and not c["isVisual"]
    synth.write_config(p, "a", "b", crlf=False)
    assert host.read_config(p)["FolderFilePath1"] == "a"
    synth.write_config(p, "a", "b", res=0.005, sv=0.5)              # SVsize > 40 * PCres (src/CommonFunc.cpp:76-79)
    assert host.read_config(p) is None
    synth.write_config(p, "a", "b", dtinit=0.001, dtmin=0.004)      # DTinit < DTmin (:120-123)
    assert host.read_config(p) is None
    synth.write_config(p, "a", "b", res=0.0)                        # PCres <= 0 (:52-55)
    assert host.read_config(p) is None
    assert host.read_config(str(tmp_path / "missing.txt")) is None


def test_pcd_roundtrip_and_epoch_listing(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(0, 5, (1000, 3)).astype(np.float32)
    for e in (3, 1, 12, 2):
        synth.write_pcd(str(tmp_path / ("Epoch_%03d.pcd" % e)), xyz + e)
    assert host.list_epochs(str(tmp_path)) == [1, 2, 3, 12]
    assert host.list_epochs(str(tmp_path) + "/") == [1, 2, 3, 12]
    assert np.array_equal(host.load_pcd(str(tmp_path / "Epoch_003.pcd")), xyz + 3)
    assert host.save_pcd(str(tmp_path / "out.pcd"), xyz)
    assert np.array_equal(host.load_pcd(str(tmp_path / "out.pcd")), xyz)
    # ascii PCD with an extra field
    with open(tmp_path / "a.pcd", "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA ascii\n"
                "1.5 2.5 3.5 9\n-1 -2 -3 8\n")
    assert np.array_equal(host.load_pcd(str(tmp_path / "a.pcd")), np.array([[1.5, 2.5, 3.5], [-1, -2, -3]], np.float32))


def test_patch_normal_matches_oracle(oracle):
    rng = np.random.default_rng(9)
    for _ in range(40):
        n = rng.normal(0, 1, 3); n /= np.linalg.norm(n)
        u = np.cross(n, [1, 0, 0]); u /= np.linalg.norm(u); v = np.cross(n, u)
        pts = (rng.uniform(-3, 3, 3) + rng.uniform(-0.03, 0.03, (40, 1)) * u + rng.uniform(-0.03, 0.03, (40, 1)) * v
               + rng.normal(0, 5e-4, (40, 1)) * n).astype(np.float32)
        a, oka = host.patch_normal(pts)
        b, okb = oracle.patch_normal(pts)
        assert oka == okb and np.array_equal(a, b)
    a, ok = host.patch_normal(np.zeros((3, 3), np.float32))
    assert not ok and np.array_equal(a, [0, 0, 1])


def test_patch_generation_stand_in():
    pts = synth.make_scan(extent=1.5, spacing=0.01, seed=5)
    p = host.patches(pts, 0.1)
    n = len(p["ct"])
    assert 100 < n <= 15 * 15 + 40 and p["bp"].shape == (6 * n, 3)
    assert (p["bpstd"] > 0).all() and (p["bpstd"] < 2e-3).all() and (p["ctstd"] < p["bpstd"]).all()
    bp = p["bp"].reshape(n, 6, 3)
    assert (bp[:, 0, 0] >= bp[:, 1, 0]).all() and (bp[:, 2, 1] >= bp[:, 3, 1]).all() and (bp[:, 4, 2] >= bp[:, 5, 2]).all()
    assert (np.abs(bp - p["ct"][:, None, :]).max(axis=(1, 2)) < 0.11).all()


def test_segmenter_plugin_groups_points_like_the_reference():
    """pwicp_host_set_segmenter: labels from the caller's segmentation replace the cubic-cell stand-in; points are grouped per
    label in cloud order (src/Segmentation.cpp:95-100) and go through the reference's refinement / planarity gates."""
    import ctypes as C
    pts = synth.make_scan(extent=1.0, spacing=0.01, seed=6)
    calls = {}

    @C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int))
    def stripes(xyz, n, sv, knn, labels):
        a = np.ctypeslib.as_array(xyz, shape=(n, 3))
        lab = np.floor((a[:, 0] - a[:, 0].min()) / sv).astype(np.int32) * 64 + np.floor((a[:, 1] - a[:, 1].min()) / sv).astype(np.int32)
        uniq, inv = np.unique(lab, return_inverse=True)
        np.ctypeslib.as_array(labels, shape=(n,))[:] = inv
        calls["n"], calls["knn"], calls["sv"], calls["labels"] = n, knn, sv, inv.copy()
        return len(uniq)

    L = host.lib()
    L.pwicp_host_set_segmenter.argtypes = [C.c_void_p]
    L.pwicp_host_set_segmenter(C.cast(stripes, C.c_void_p))
    try:
        p = host.patches(pts, 0.1)
    finally:
        L.pwicp_host_set_segmenter(None)
    assert calls["n"] == len(pts) and calls["knn"] == 45 and abs(calls["sv"] - 0.1) < 1e-7
    n = len(p["ct"])
    assert 50 < n <= calls["labels"].max() + 1
    # every centroid is the mean of a refined subset of one label's points: it lies inside that label's bounding box
    lab = calls["labels"]
    for k in range(0, n, 7):
        d = np.abs(pts - p["ct"][k]).max(axis=1)
        owner = lab[np.argmin(d)]
        box = pts[lab == owner]
        assert (p["ct"][k] >= box.min(0) - 1e-6).all() and (p["ct"][k] <= box.max(0) + 1e-6).all()
    # a failing segmenter is fatal in the reference-shaped mirror only through its return code: here it just restores
    q = host.patches(pts, 0.1)                       # stand-in again
    assert len(q["ct"]) > 0 and len(q["ct"]) != n or not np.array_equal(q["ct"], p["ct"])


def test_segmenter_plugin_by_environment(oracle, tmp_path):
    """PWICP_SEGMENTER_PLUGIN=<so>:<symbol> (include/pwicp_host.h): with the reference's own segmentation compiled where it
    lies (oracle/_ref, test side only) the mirror's patch post-processing selects the patches the reference selected on its
    shipped pair -- same centroids, same boundary points (tests/golden/refpair_e2.npz holds them).  Runs in a child process:
    the plug-in is looked up once per process."""
    import subprocess, sys
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_supervoxel.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref/libref_supervoxel.so not built (needs /root/reference at build time)")
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests')); sys.path.insert(0, os.path.join(%r, 'piecewise-icp_b200', 'python'))\n"
        "from conftest import load_refpair\n"
        "from oracle import oracle_py as O\n"
        "from pwicp_b200 import host\n"
        "f = load_refpair(O.patch_stats)\n"
        "for cloud, ct_ref, bp_ref in ((f['pair']['cloud1'], f['pair']['ct1'], None), (f['pair']['cloud2'], f['pair']['ct2'], f['pair']['bp2'])):\n"
        "    p = host.patches(cloud, f['pair']['SVRes1'])\n"
        "    assert p['ct'].shape == ct_ref.shape and np.array_equal(p['ct'], ct_ref)\n"
        "    assert bp_ref is None or np.array_equal(p['bp'], bp_ref)\n"
        "print('PLUGIN-OK')\n") % (ROOT, ROOT, ROOT)
    env = dict(os.environ, PWICP_SEGMENTER_PLUGIN=ref_so + ":ref_supervoxel_labels")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "PLUGIN-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _write_transmatrices(path, Ts, Vs):
    with open(path, "w") as f:
        for k, (T, V) in enumerate(zip(Ts, Vs)):
            f.write("%d\n" % (k + 2))
            for r in range(4): f.write(" ".join("%.12f" % v for v in T[r]) + " \n")
            for r in range(6): f.write(" ".join("%.12f" % v for v in V[r]) + " \n")


def _read_blocks(path, n):
    vals = open(path).read().split()
    out, p = [], 0
    for _ in range(n):
        t = int(vals[p]); p += 1
        T = np.array(vals[p:p + 16], float).reshape(4, 4); p += 16
        V = np.array(vals[p:p + 36], float).reshape(6, 6); p += 36
        out.append((t, T, V))
    return out


def test_chain_to_reference_epoch(tmp_path):
    """calTransToReferenceEpoch (src/Registration.cpp:977-1153): direct, fixed-interval and adaptive."""
    rng = np.random.default_rng(2)
    n = 5
    Ts = [synth.rigid_matrix(*rng.uniform(-0.01, 0.01, 6)) for _ in range(n)]
    Vs = [np.diag(rng.uniform(1e-10, 1e-8, 6)) for _ in range(n)]
    tm = str(tmp_path / "TransMatrices.txt")
    _write_transmatrices(tm, Ts, Vs)
    f32 = lambda M: M.astype(np.float32).astype(np.float64)
    # mode 0: identity chain
    host.lib().pwicp_host_chain_to_reference(tm.encode(), 0, b"", n, str(tmp_path / "o0.txt").encode(), str(tmp_path / "p0.txt").encode())
    for k, (t, T, V) in enumerate(_read_blocks(str(tmp_path / "o0.txt"), n)):
        assert t == k + 2 and np.allclose(T, Ts[k], atol=1e-6) and np.allclose(V, Vs[k], rtol=1e-3)
    # mode 2: T_i * T_{i-2} * ..., VCMs added
    host.lib().pwicp_host_chain_to_reference(tm.encode(), 2, b"", n, str(tmp_path / "o2.txt").encode(), str(tmp_path / "p2.txt").encode())
    blocks = _read_blocks(str(tmp_path / "o2.txt"), n)
    for i in range(n):
        T, V, j = np.eye(4), np.zeros((6, 6)), i
        while True:
            T = Ts[j] @ T; V = V + Vs[j]
            if j < 2: break
            j -= 2
        assert np.allclose(blocks[i][1], T, atol=2e-6) and np.allclose(blocks[i][2], V, rtol=1e-3)
    # adaptive: pairs 1->0, 2->1, 3->1, 4->3, 5->4 with the adjoint propagation
    pairs = {1: 0, 2: 1, 3: 1, 4: 3, 5: 4}
    pf = str(tmp_path / "RegPairFile.txt")
    open(pf, "w").write("".join(f"{s} {t}\n" for s, t in pairs.items()))
    host.lib().pwicp_host_chain_to_reference(tm.encode(), -1, pf.encode(), n, str(tmp_path / "oa.txt").encode(), str(tmp_path / "pa.txt").encode())
    blocks = _read_blocks(str(tmp_path / "oa.txt"), n)
    for i in range(n):
        T, V, tgt = f32(Ts[i]), Vs[i].copy(), i + 1
        while True:
            tgt = pairs[tgt]
            if tgt == 0: break
            M = f32(Ts[tgt - 1]); R, t = M[:3, :3], M[:3, 3]
            S = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
            Ad = np.zeros((6, 6)); Ad[:3, :3] = R; Ad[3:, 3:] = R; Ad[3:, :3] = S @ R
            T = M @ T; V = Vs[tgt - 1] + Ad @ V @ Ad.T
        assert np.allclose(blocks[i][1], T, atol=2e-6) and np.allclose(blocks[i][2], V, rtol=2e-3, atol=2e-12)  # files carry 12 decimals
    hdr = open(tmp_path / "pa.txt").readline()
    assert hdr.startswith("Epoch  Rx[gon]  Ry[gon]  Rz[gon]  tx[m]  ty[m]  tz[m]  Std_Rx[mgon]")


def test_chain_and_error_report_reproduce_the_references_recorded_files(tmp_path):
    """F2 and A10 against the reference's own recorded outputs (tests/golden/recorded_4d, copied from results/4DPCReg and
    data/data_synthetic): from the recorded TransMatrices.txt (adaptive run) and the adaptive pair list,
    calTransToReferenceEpoch (src/Registration.cpp:977-1153, adjoint covariance propagation :1072-1083) must write the
    recorded TransMatrices_toRef.txt / TransParameters_toRef.txt, and calAbsErrorOfTransPara (:1157-1251, matrix2angle
    src/CommonFunc.cpp:385-407) the recorded TransPara_AbsError.txt -- text for text."""
    import ctypes as C
    rec = os.path.join(ROOT, "tests", "golden", "recorded_4d")
    norm = lambda path: open(path).read().replace("\r", "").rstrip("\n")
    L = host.lib()
    out_tm, out_tp = str(tmp_path / "TransMatrices_toRef.txt"), str(tmp_path / "TransParameters_toRef.txt")
    L.pwicp_host_chain_to_reference(os.path.join(rec, "TransMatrices.txt").encode(), -1, os.path.join(rec, "RegPairFile.txt").encode(), 19,
                                    out_tm.encode(), out_tp.encode())
    assert norm(out_tm) == norm(os.path.join(rec, "TransMatrices_toRef.txt"))
    assert norm(out_tp) == norm(os.path.join(rec, "TransParameters_toRef.txt"))
    L.pwicp_host_abs_error.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    L.pwicp_host_abs_error.restype = None
    out_err = str(tmp_path / "TransPara_AbsError.txt")
    L.pwicp_host_abs_error(os.path.join(rec, "TransMatrices_toRef.txt").encode(), os.path.join(rec, "defined_transformations.txt").encode(), 20, 0,
                           out_err.encode())
    assert norm(out_err) == norm(os.path.join(rec, "TransPara_AbsError.txt"))
    # the per-pair result file: layout, 12 / 10 decimals, gon angles of matrix2angle -- line for line down to the VCM; the six
    # standard deviations only to the 4 digits the printed VCM still carries
    L.pwicp_host_write_transmatrix.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    for f in ("2_Direct2Ref_TransMatrix.txt", "12_Adaptive_TransMatrix.txt"):
        want = norm(os.path.join(rec, f)).splitlines()
        T = np.array([[float(v) for v in want[1 + r].split()] for r in range(4)], np.float32)
        V = np.array([[float(v) for v in want[16 + r].split()] for r in range(6)])
        assert L.pwicp_host_write_transmatrix(str(tmp_path / f).encode(), T.ctypes.data, V.ctypes.data)
        got = norm(str(tmp_path / f)).splitlines()
        assert len(got) == len(want) == 30 and got[:24] == want[:24]
        for a, b in zip(got[24:], want[24:]):
            assert a.split()[:2] == b.split()[:2] and a.split()[-1] == b.split()[-1]
            assert abs(float(a.split()[2]) / float(b.split()[2]) - 1) < 2e-3


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "piecewise-icp_b200", "python"))
import numpy as np, torch.distributed as dist
from pwicp_b200 import host
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 7
recs = (host.EpochRecord * n)()
for k in range(n):
    recs[k].step = k + 1
    if k % world == rank:            # epoch sharding: (step - 1) % world == rank
        recs[k].status = 1; recs[k].time_stamp = 100 + k
        for j in range(16): recs[k].T[j] = rank * 1000 + k + j / 100.0
        recs[k].VCM[0] = 1e-9 * (k + 1)
merged = host.gather_records(recs, dist)
out = host.array_to_records(merged)
ok = all(out[k].status == 1 and out[k].time_stamp == 100 + k and abs(out[k].T[3] - ((k % world) * 1000 + k + 0.03)) < 1e-3
         and abs(out[k].VCM[0] - 1e-9 * (k + 1)) < 1e-15 for k in range(n))
print("RANK", rank, "OK" if ok else "BAD")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_epoch_record_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2
