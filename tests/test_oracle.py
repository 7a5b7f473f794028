"""CPU tests of the oracle (no GPU): independent cross-checks + the committed golden vectors.

The oracle is pinned by the reference: its recorded results on its shipped scans (the committed
pair tests/golden/refpair_e2.npz here; all recorded pairs when /root/reference is present), its own KD-tree compiled into
oracle/_ref, and its recorded aggregate files.  On top of that it is checked against independent implementations (brute
force, scipy, numpy float64 algebra) and physical properties.
"""
import os

import numpy as np
import pytest

from pwicp_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pair2k.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def pair2k():
    return synth.make_pair(2000, seed=20250606)


def test_nn_matches_brute_force_and_scipy(oracle):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    for n1, nq in [(1, 10), (7, 50), (16, 40), (500, 700), (5000, 3000)]:
        tgt = rng.normal(0, 1, (n1, 3)).astype(np.float32)
        qry = rng.normal(0, 1.5, (nq, 3)).astype(np.float32)
        i1, d1 = oracle.nn(tgt, qry)
        i2, d2 = oracle.nn(tgt, qry, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        dd, ii = cKDTree(tgt.astype(np.float64)).query(qry.astype(np.float64))
        assert (ii == i1).mean() > 0.999          # float32 vs float64 may differ on near ties
        assert np.allclose(np.sqrt(d1), dd, rtol=1e-5, atol=1e-7)


def _assert_same_nn(tgt, qry, idx_a, idx_ref):
    """Indices must agree except on rounding-level ties: the reference KD-tree's metric is accumulated in double
    (codelibrary/util/metric/squared_euclidean.h:33-36), the hot path's in float (flann::L2_Simple<float>)."""
    bad = np.nonzero(idx_a != idx_ref)[0]
    if len(bad):
        def f32d2(i):
            d = qry[bad] - tgt[i[bad]]
            return ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
        a, r = f32d2(idx_a), f32d2(idx_ref)           # idx_a is the float minimum; the other within 2 ulp of it
        assert (r >= a).all() and (r <= a * np.float32(1 + 3e-7)).all(), f"{len(bad)} real NN mismatches"
    return len(bad)


def test_nn_pinned_by_reference_kdtree(oracle, gold, pair2k):
    """oracle/_ref: the reference's own KD-tree (codelibrary/util/tree/kd_tree.h, the structure FLANN's
    KDTreeSingleIndex shares) compiled from /root/reference where it lies.  The oracle's exact 1-NN and the
    committed golden index vectors must be what reference-authored code returns."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_kdtree.so not built (no /root/reference on this machine)")
    # (i) the golden vectors
    q = np.concatenate([pair2k["ct2"], pair2k["bp2"]])
    ri, rd = oracle.ref_nn(pair2k["ct1"], q)
    assert _assert_same_nn(pair2k["ct1"], q, gold["nn_pair_idx"], ri) == 0
    rng = np.random.default_rng(7)
    tgt = rng.uniform(-5, 5, (100000, 3)).astype(np.float32)
    qry = rng.uniform(-6, 6, (20000, 3)).astype(np.float32)
    ri, rd = oracle.ref_nn(tgt, qry)
    assert _assert_same_nn(tgt, qry, gold["nn_rand_idx"], ri) == 0
    assert np.allclose(gold["nn_rand_d2"], rd, rtol=3e-7)
    # (ii) the oracle itself on a fresh 60k centroid pair (420k queries incl. boundary points) and on clustered data
    d = synth.make_pair(60000, seed=11)
    q = np.concatenate([d["ct2"], d["bp2"]])
    oi, od = oracle.nn(d["ct1"], q)
    ri, rd = oracle.ref_nn(d["ct1"], q)
    _assert_same_nn(d["ct1"], q, oi, ri)
    assert np.allclose(od, rd, rtol=3e-7, atol=1e-12)
    rng = np.random.default_rng(3)
    centres = rng.normal(0, 10, (50, 3))
    tgt = (centres[rng.integers(0, 50, 30000)] + rng.normal(0, 0.05, (30000, 3))).astype(np.float32)
    qry = (centres[rng.integers(0, 50, 20000)] + rng.normal(0, 0.3, (20000, 3))).astype(np.float32)
    oi, _ = oracle.nn(tgt, qry)
    ri, _ = oracle.ref_nn(tgt, qry)
    _assert_same_nn(tgt, qry, oi, ri)
    # (iii) the k = 2 self search behind calPCresolution (src/CommonFunc.cpp:239-263): neighbour 0 is the point itself
    kn = oracle.ref_knn(tgt[:5000], tgt[:5000], 2)
    oi2, _ = oracle.nn(tgt[:5000], tgt[:5000])
    assert np.array_equal(kn[:, 0], np.arange(5000)) or (np.linalg.norm(tgt[kn[:, 0]] - tgt[:5000], axis=1) == 0).all()
    assert (np.linalg.norm(tgt[oi2] - tgt[:5000], axis=1) == 0).all()


def test_nn_ties_resolve_to_lowest_index(oracle):
    tgt = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [1, 0, 0]], np.float32)
    idx, d2 = oracle.nn(tgt, np.zeros((1, 3), np.float32))
    assert idx[0] == 0 and d2[0] == 1.0
    big = np.tile(tgt, (40, 1))                   # duplicates spread over several leaves
    idx, _ = oracle.nn(big, np.zeros((3, 3), np.float32))
    assert (idx == 0).all()


def test_nn_distance_is_float32_left_to_right(oracle):
    rng = np.random.default_rng(1)
    tgt = rng.normal(0, 3, (200, 3)).astype(np.float32)
    qry = rng.normal(0, 3, (100, 3)).astype(np.float32)
    idx, d2 = oracle.nn(tgt, qry)
    d = qry - tgt[idx]
    expect = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    assert np.array_equal(d2, expect.astype(np.float32))


def test_transform_is_float32_in_order(oracle):
    rng = np.random.default_rng(2)
    p = rng.normal(0, 10, (1000, 3)).astype(np.float32)
    T = synth.rigid_matrix(0.1, -0.2, 0.3, 1, 2, 3).astype(np.float32)
    out = oracle.transform(p, T)
    exp = np.empty_like(p)
    for r in range(3):
        exp[:, r] = ((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]
    assert np.array_equal(out, exp)


def test_lls_step_matches_float64_least_squares(oracle, pair2k):
    d = pair2k
    idx, _ = oracle.nn(d["ct1"], d["ct2"])
    ATA, ATb, x, T = oracle.lls_step(d["ct2"], idx, d["ct1"], d["nrm1"])
    s, q, n = d["ct2"].astype(np.float64), d["ct1"][idx].astype(np.float64), d["nrm1"][idx].astype(np.float64)
    A = np.concatenate([np.cross(s, n), n], 1)
    b = ((q - s) * n).sum(1)
    assert np.allclose(ATA, A.T @ A, rtol=1e-5)
    xs = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.allclose(x, xs, rtol=1e-3, atol=1e-7)
    assert np.allclose(ATA, ATA.T)
    assert T.dtype == np.float32 and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)


def test_gpu_summation_order_is_equivalent(oracle, pair2k):
    d = pair2k
    idx, _ = oracle.nn(d["ct1"], d["ct2"])
    a = oracle.lls_step(d["ct2"], idx, d["ct1"], d["nrm1"], 0)
    for gb in (2, 7, 32, 1000):
        b = oracle.lls_step(d["ct2"], idx, d["ct1"], d["nrm1"], 1, gb)
        assert np.allclose(a[0], b[0], rtol=1e-12) and np.allclose(a[2], b[2], rtol=1e-9, atol=1e-15)
        assert np.abs(a[3] - b[3]).max() <= 6e-8      # one float ulp at most


def test_icp_recovers_a_known_rigid_motion(oracle):
    d = synth.make_pair(2500, changed=0.0, motion=(0.002, -0.001, 0.0015, 0.003, -0.002, 0.001))
    T = np.eye(4, dtype=np.float32)
    src = d["ct2"].copy()
    for _ in range(6):                                # the reference restarts ICP every outer iteration
        r = oracle.icp(d["ct1"], d["nrm1"], src)
        src = oracle.transform(src, r["T"])
        T = oracle.mat4_mul(r["T"], T)
    err = np.abs(T.astype(np.float64) - d["T_true"]).max()
    assert err < 2e-4, err
    assert r["state"] in (2, 3, 4)


def test_icp_forced_iterations_and_states(oracle, pair2k):
    d = pair2k
    r = oracle.icp(d["ct1"], d["nrm1"], d["ct2"], oracle.icp_params(max_iter=7, force_iters=1), trace=True)
    assert r["n_iter"] == 7 and r["state"] == 1 and len(r["mse"]) == 7
    r1 = oracle.icp(d["ct1"], d["nrm1"], d["ct2"], oracle.icp_params(max_iter=1))
    assert r1["n_iter"] == 1 and r1["state"] == 1
    # final transform is the ordered product of the incremental ones (final = T * final)
    F = np.eye(4, dtype=np.float32)
    for Tk in r["T_trace"]:
        F = oracle.mat4_mul(Tk, F)
    assert np.array_equal(F, r["T"])


def test_octree_bbox_is_a_padded_cube(oracle):
    rng = np.random.default_rng(5)
    for scale in [(10, 4, 1), (0.3, 0.3, 0.3), (50, 2, 7)]:
        p = (rng.uniform(-1, 1, (3000, 3)) * scale).astype(np.float32)
        for res in (0.01, 0.25, 2.0):
            bb = oracle.octree_bbox(p, res)
            side = bb[3:] - bb[:3]
            assert np.allclose(side, side[0], rtol=1e-9)
            k = np.log2(side[0] / res)
            assert abs(k - round(k)) < 1e-6 and side[0] >= 2 * res - 1e-9
            assert (p >= bb[:3] - 1e-6).all() and (p <= bb[3:] + 1e-6).all()
            c = (p.max(0).astype(np.float64) + p.min(0)) / 2
            assert np.allclose((bb[:3] + bb[3:]) / 2, c, atol=1e-3)


def test_percentile_matches_numpy(oracle, pair2k):
    d = pair2k
    v = oracle.percentile_nn(d["cloud1"], d["cloud2"], 0.75)
    _, d2 = oracle.nn(d["cloud1"], d["cloud2"])
    s = np.sort(np.sqrt(d2).astype(np.float64))
    assert v == s[int(np.float32(len(s)) * np.float32(0.75))]


def test_vcm_matches_numpy(oracle, pair2k):
    d = pair2k
    src = d["ct2"][~d["changed"]][:400]
    V, sing = oracle.vcm(d["ct1"], d["nrm1"], src)
    idx, _ = oracle.nn(d["ct1"], src)
    Q, P, N = src.astype(np.float64), d["ct1"][idx].astype(np.float64), d["nrm1"][idx].astype(np.float64)
    A = np.concatenate([np.cross(Q, N), N], 1)
    L = (N * (P - Q)).sum(1)
    Qxx = np.linalg.inv(A.T @ A)
    X = Qxx @ A.T @ L
    v = A @ X - L
    D = (v @ v) / (len(src) - 6) * Qxx
    assert np.allclose(V, D, rtol=1e-6, atol=0) and not sing
    assert np.allclose(V, V.T, rtol=1e-9)


def test_patch_normal_matches_eigh(oracle):
    rng = np.random.default_rng(9)
    for _ in range(50):
        n = rng.normal(0, 1, 3); n /= np.linalg.norm(n)
        u = np.cross(n, [1, 0, 0]); u /= np.linalg.norm(u); v = np.cross(n, u)
        c = rng.uniform(-3, 3, 3)
        ab = rng.uniform(-0.03, 0.03, (40, 2))
        pts = (c + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, 5e-4, (40, 1)) * n).astype(np.float32)
        est, ok = oracle.patch_normal(pts)
        assert ok and abs(np.linalg.norm(est) - 1) < 1e-5
        w, V = np.linalg.eigh(np.cov(pts.astype(np.float64).T))
        assert abs(abs(est @ V[:, 0]) - 1) < 2e-3
    est, ok = oracle.patch_normal(np.zeros((4, 3), np.float32))     # <= 4 points -> (0,0,1), failure
    assert not ok and np.array_equal(est, [0, 0, 1])


def test_matrix2angle_roundtrip(oracle):
    rng = np.random.default_rng(4)
    for _ in range(50):
        a = rng.uniform(-1.2, 1.2, 3)
        T = synth.rigid_matrix(*a, 0, 0, 0).astype(np.float32)
        assert np.allclose(oracle.matrix2angle(T), a, atol=2e-6)


def test_outer_loop_properties(oracle, pair2k):
    d = pair2k
    pd = oracle.PairData(d)
    res = oracle.piecewise_icp(pd, 1, 0.05)
    s = res["DTseries"]
    assert res["rc"] > 0 and len(s) == res["rc"] + 1
    assert (np.diff(s) <= 0).all()                       # monotonically decreasing DT (:907)
    assert s[-1] == np.float32(d["DTmin"])               # ends at LoDet_min = DTmin
    assert np.abs(res["T"].astype(np.float64) - d["T_true"]).max() < 2e-4
    # changed patches are rejected, unchanged ones kept
    st = oracle.State(0.004, 0, 0, 1, 0)
    rc, T, V, flags, stats = oracle.single_iteration(pd, st)
    assert rc == 0 and flags[d["changed"]].mean() < 0.05 and flags[~d["changed"]].mean() > 0.9
    assert np.sqrt(np.diag(res["VCM"])).max() < 1e-3


def test_outer_loop_error_codes(oracle, pair2k):
    d = dict(pair2k)
    few = {k: (v[:3] if k in ("ct2", "bpstd2") else v) for k, v in d.items()}
    few["bp2"] = d["bp2"][:18]; few["patch_off2"] = d["patch_off2"][:4]
    rc, *_ = oracle.single_iteration(oracle.PairData(few), oracle.State(0.05, 0, 0, 0, 0))
    assert rc == -1                                       # < 4 patches (:728-731)
    far = dict(d); far["ct2"] = d["ct2"] + np.float32(5.0)
    rc, *_ = oracle.single_iteration(oracle.PairData(far), oracle.State(0.05, 0, 0, 0, 0))
    assert rc == -2                                       # < 4 stable patches (:864-867)


def test_golden_vectors(oracle, gold, pair2k):
    d = pair2k
    q = np.concatenate([d["ct2"], d["bp2"]])
    i, d2 = oracle.nn(d["ct1"], q)
    assert np.array_equal(i, gold["nn_pair_idx"]) and np.array_equal(d2, gold["nn_pair_d2"])
    r = oracle.icp(d["ct1"], d["nrm1"], d["ct2"], oracle.icp_params(max_iter=10, force_iters=1), trace=True)
    assert np.array_equal(r["T_trace"], gold["icp_T_trace"]) and np.array_equal(r["mse"], gold["icp_mse"])
    res = oracle.piecewise_icp(oracle.PairData(d), 1, 0.05)
    assert np.array_equal(res["DTseries"], gold["outer_DTseries"])
    assert np.array_equal(res["T"], gold["outer_T"])
    assert np.allclose(res["VCM"], gold["outer_VCM"], rtol=1e-12)
    res2 = oracle.piecewise_icp(oracle.PairData(d), 0, 0.0)
    assert np.array_equal(res2["DTseries"], gold["outer_auto_DTseries"])


def test_patch_stats_match_numpy(oracle):
    """F3: centroid / boundary points / sigma of every patch against float64 numpy; the normal is the
    reference's single-pass float covariance (loses digits by design), so it is only checked loosely."""
    d = synth.make_pair(1500, pts_per_patch=24)
    r = oracle.patch_stats(d["patch_pts2"], d["patch_off2"])
    pp = d["patch_pts2"].reshape(-1, 24, 3)
    assert np.abs(pp.astype(np.float64).mean(1) - r["ct"]).max() < 1e-6
    for k, (axis, fn) in enumerate([(0, np.argmax), (0, np.argmin), (1, np.argmax), (1, np.argmin), (2, np.argmax), (2, np.argmin)]):
        pick = fn(pp[:, :, axis], axis=1)                       # first extremal point wins, like the strict comparisons
        assert np.array_equal(r["bp"][:, k], pp[np.arange(len(pp)), pick])
    c = pp.astype(np.float64) - pp.astype(np.float64).mean(1, keepdims=True)
    w, v = np.linalg.eigh(np.einsum("nki,nkj->nij", c, c))
    nn = v[:, :, 0]
    sd = np.sqrt((np.einsum("nki,ni->nk", c, nn) ** 2).sum(1) / 23)
    assert np.allclose(r["bpstd"], sd, rtol=1e-5) and np.allclose(r["ctstd"], sd / 24, rtol=1e-5)
    assert r["nrm_ok"].all()
    assert np.minimum(np.abs(nn - r["nrm"]).max(1), np.abs(nn + r["nrm"]).max(1)).max() < 1e-2
    # ragged patches incl. too-small ones: calPatchNormal refuses 4 points or fewer
    off = np.array([0, 3, 7, 12, 40], np.int32)
    r = oracle.patch_stats(d["patch_pts2"][:40], off)
    assert r["nrm_ok"].tolist() == [0, 0, 1, 1] and np.array_equal(r["nrm"][0], [0, 0, 1])


# ---------------------------------------------------------------- pinned by the reference's own recorded result
def test_outer_loop_reproduces_the_references_recorded_result(oracle):
    """BASELINE configs[0].  The reference ships the scans (data/data_synthetic), the configuration
    (configuration_files/configuration_4d.txt) and the result its own build wrote for them
    (results/4DPCReg/2_Direct2Ref_TransMatrix.txt).  tests/golden/refpair_e2.npz holds that pair at the hot-path boundary,
    segmented by the reference's own supervoxel code (oracle/_ref, tests/golden/make_refpair.py).  The oracle's outer loop
    on it must land on the recorded 4x4 within the north-star tolerance (1e-6 rad / 1e-6 m; the file prints 12 decimals
    of float32 values) and on the recorded VCM."""
    from conftest import load_refpair
    f = load_refpair(oracle.patch_stats)
    d = f["pair"]
    assert len(d["ct1"]) == 1822 and len(d["ct2"]) == 1846
    res = oracle.piecewise_icp(oracle.PairData(d), 1, f["DTinit"])
    T = oracle.mat4_mul(oracle.mat4_mul(f["Sinv"], res["T"]), f["S"])          # src/Registration.cpp:461
    da = np.abs(oracle.matrix2angle(T) - oracle.matrix2angle(f["T_recorded"].astype(np.float32))).max()
    dt = np.abs(T[:3, 3] - f["T_recorded"][:3, 3]).max()
    assert da <= 1e-6 and dt <= 1e-6, (da, dt)
    assert np.abs(T - f["T_recorded"]).max() <= 1e-6
    sd, sd_rec = np.sqrt(np.diag(res["VCM"])), np.sqrt(np.diag(f["VCM_recorded"]))
    assert np.allclose(sd, sd_rec, rtol=2e-3)                                   # recorded with ~3-4 significant digits
    assert np.allclose(res["VCM"], f["VCM_recorded"], rtol=0, atol=2e-3 * np.abs(f["VCM_recorded"]).max())
    assert len(res["DTseries"]) - 1 == 4 and (np.diff(res["DTseries"]) <= 0).all()
    # and it is a good registration: the recorded result's own distance to the ground truth (1.1 mm, 1e-4 rad)
    G = f["T_truth"]
    e = min(np.abs(T - X).max() for X in (G, np.linalg.inv(G)))
    assert e < 1.5e-3


# ---------------------------------------------------------------- F4: VoxelGrid / StatisticalOutlierRemoval restatement
def test_preprocessing_restatement(oracle):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(8)
    c = rng.uniform(-1, 1, (20000, 3)).astype(np.float32)
    v = oracle.voxel_grid(c, 0.1)
    key = np.floor(c * np.float32(1.0 / np.float32(0.1))).astype(np.int64)
    key -= key.min(0)
    dims = key.max(0) + 1
    lin = key[:, 0] + key[:, 1] * dims[0] + key[:, 2] * dims[0] * dims[1]
    uniq = np.unique(lin)
    assert len(v) == len(uniq)
    ref = np.stack([c[lin == u].astype(np.float64).mean(0) for u in uniq[:200]])       # ascending voxel index
    assert np.allclose(v[:200], ref, atol=1e-6)
    md = oracle.knn_mean_dist(c, 14)
    dd, _ = cKDTree(c.astype(np.float64)).query(c.astype(np.float64), 15)
    assert np.allclose(md, dd[:, 1:].mean(1), rtol=1e-5)
    out, thr = oracle.sor_select(c, md, 1.0)
    assert np.isclose(thr, md.astype(np.float64).mean() + md.astype(np.float64).std(ddof=1), rtol=1e-9)
    assert np.array_equal(out, c[md <= thr])
    # the host mirror's statements of the same filters give the identical cloud
    from pwicp_b200 import host
    scan = synth.make_scan(extent=1.0, spacing=0.005, seed=4)
    assert np.array_equal(host.preprocess(scan, 0.005, 14, 5.0, device=False), oracle.preprocess(scan, 0.005, 14, 5.0))
    # within-voxel order of the Microsoft STL's std::sort (the reference's Windows build): same voxels, same points up to the
    # last bit of some centroids; host statements (PWICP_VOXEL_ORDER=msvc) and oracle agree bit for bit
    dense = (scan[:, None, :] + rng.normal(0, 4e-4, (len(scan), 4, 3))).reshape(-1, 3).astype(np.float32)    # ~4 points per voxel
    a, b = oracle.voxel_grid(dense, 0.005), oracle.voxel_grid(dense, 0.005, msvc_order=True)
    assert a.shape == b.shape and np.allclose(a, b, atol=1e-6) and (a != b).any()
    os.environ["PWICP_VOXEL_ORDER"] = "msvc"
    try:
        assert np.array_equal(host.preprocess(dense, 0.005, 14, 5.0, device=False), oracle.preprocess(dense, 0.005, 14, 5.0, msvc_order=True))
    finally:
        del os.environ["PWICP_VOXEL_ORDER"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/data/data_synthetic"), reason="needs the reference tree (absent on the GPU box)")
def test_recorded_results_reproduced_on_more_shipped_pairs(oracle):
    """scripts/refdata_oracle.py on more of the reference's shipped pairs, all three recorded pair modes (epochs with 4 to 6 outer
    iterations, all three DT stages): host pre-processing mirror + the reference's own segmentation (oracle/_ref) + the oracle's outer loop
    against results/4DPCReg/<e>_Direct2Ref_TransMatrix.txt, within the north-star tolerance.  (All 19 pairs:
    profiles/r01i_refdata_oracle_cpu.txt -- 16 within 1e-6, the rest input-side)"""
    import sys
    if not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_supervoxel.so")):
        pytest.skip("oracle/_ref/libref_supervoxel.so not built")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import refdata_oracle as R
    from pwicp_b200 import host
    scans = os.path.join(R.REF, "data/data_synthetic/syntheticPC_with_transformations")
    load = lambda k: host.load_pcd(os.path.join(scans, "Epoch_%03d.pcd" % k))
    # (source epoch, target epoch, recorded family, outer iterations): reference-epoch mode, fixed interval 3, adaptive
    for e, tgt, tag, n_outer in ((4, 1, "Direct2Ref", 4), (12, 1, "Direct2Ref", 6), (16, 1, "Direct2Ref", 6),
                                 (12, 9, "Fixed", 6), (14, 9, "Adaptive", 5)):
        T, res, d = R.register(load(tgt), load(e))
        Tr, Vr = R.read_T(os.path.join(R.REF, "results/4DPCReg/%d_%s_TransMatrix.txt" % (e, tag)))
        da, dt = R.pose_err(T, Tr)
        assert da <= 1e-6 and dt <= 1e-6, (e, tag, da, dt)
        assert len(res["DTseries"]) - 1 == n_outer
        assert np.allclose(np.sqrt(np.diag(res["VCM"])), np.sqrt(np.diag(Vr)), rtol=2e-3)
    # the three pairs of the reference-epoch family that miss with the input order inside a voxel (2.9e-6, 1.8e-5, 8.7e-4 rad)
    # are reproduced once pcl::VoxelGrid's points are summed in the order of the Microsoft STL's std::sort (host/msvc_sort.h):
    # the recorded files come from the reference's Windows build
    os.environ["PWICP_VOXEL_ORDER"] = "msvc"
    try:
        for e in (3, 8, 19):
            T, res, d = R.register(load(1), load(e))
            Tr, Vr = R.read_T(os.path.join(R.REF, "results/4DPCReg/%d_Direct2Ref_TransMatrix.txt" % e))
            da, dt = R.pose_err(T, Tr)
            assert da <= 1e-6 and dt <= 1e-6, (e, "msvc order", da, dt)
            assert np.allclose(np.sqrt(np.diag(res["VCM"])), np.sqrt(np.diag(Vr)), rtol=2e-3)
    finally:
        del os.environ["PWICP_VOXEL_ORDER"]


def test_device_order_sums_with_a_separate_first_level_fan_in(oracle):
    """reduce_mode = 1 (the kernel's hierarchical summation order): group_batches may carry a different fan-in for the first
    level in bits 16.. (one-barrier reduction); absent or equal it is the uniform hierarchy."""
    d = synth.make_pair(40000, seed=3)
    idx, _ = oracle.nn(d["ct1"], d["ct2"])
    args = (d["ct2"], idx, d["ct1"], d["nrm1"])
    base = oracle.lls_step(*args, reduce_mode=1, group_batches=32)
    assert np.array_equal(base[0], oracle.lls_step(*args, reduce_mode=1, group_batches=32 | (32 << 16))[0])
    two = oracle.lls_step(*args, reduce_mode=1, group_batches=32 | (8 << 16))
    seq = oracle.lls_step(*args, reduce_mode=0)
    assert not np.array_equal(two[0], base[0])                       # another order ...
    assert np.allclose(two[0], seq[0], rtol=1e-12) and np.allclose(two[1], seq[1], rtol=1e-10, atol=1e-18)   # ... of the same sums
    assert np.abs(two[3] - seq[3]).max() <= 1e-7


def test_knn_normals_restatement_matches_the_references_code(oracle):
    """Groundwork for the next F4 kernel (kNN-45 PCA normals, src/Segmentation.cpp:28-46): the oracle's restatement against the
    reference's own codelibrary (oracle/_ref) on the reference pair's target cloud -- identical neighbour lists, bit-identical
    normals (the closed-form smallest eigenvector of pca_estimate_normals.h)."""
    if not oracle.ref_available() or not os.path.exists(os.path.join(os.path.dirname(oracle.REF_KDTREE), "libref_supervoxel.so")):
        pytest.skip("oracle/_ref not built")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpair_e2.npz"))
    c = z["cloud1"][:40000]
    nb, nr = oracle.knn_normals(c, 45)
    rnb, rnr = oracle.ref_knn_normals(c, 45)
    assert (nb[:, 0] == np.arange(len(c))).all()                       # the point itself first
    same = (nb == rnb).all(axis=1)
    assert same.mean() > 0.999                                         # exact ties may be visited in another order ...
    for a, b in zip(nb[~same], rnb[~same]):
        assert set(a) == set(b)                                        # ... but the set is the same
    assert np.array_equal(nr[same], rnr[same])
    assert np.allclose(np.linalg.norm(nr, axis=1), 1.0, atol=1e-12)


# ---------------------------------------------------------------- property tests (hypothesis)
def test_properties_nn_and_knn_against_brute_force(oracle):
    """Random small clouds incl. duplicates, lattice points (exact ties) and far queries: the KD-tree search equals brute force
    bit for bit (index and float distance, lowest index on ties); the k-NN mean distance equals a numpy restatement."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    coords = st.integers(min_value=-6, max_value=6)

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(st.lists(st.tuples(coords, coords, coords), min_size=1, max_size=60),
           st.lists(st.tuples(coords, coords, coords), min_size=1, max_size=30), st.floats(0.01, 3.0), st.integers(0, 2**31 - 1))
    def check(tgt, qry, scale, seed):
        rng = np.random.default_rng(seed)
        t = (np.array(tgt, np.float32) * np.float32(scale)).astype(np.float32)
        q = (np.array(qry, np.float32) * np.float32(scale) + rng.choice([0.0, 0.0, 0.37], (len(qry), 3))).astype(np.float32)
        i1, d1 = oracle.nn(t, q)
        i2, d2 = oracle.nn(t, q, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
        d = q[:, None, :] - t[None, :, :]
        ref = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).astype(np.float32)
        assert np.array_equal(d1, ref.min(axis=1)) and np.array_equal(i1, ref.argmin(axis=1))        # argmin = lowest index on ties
        k = min(5, len(t) - 1)
        if k >= 1:
            md = oracle.knn_mean_dist(t, k)
            dd = t[:, None, :] - t[None, :, :]
            full = ((dd[..., 0] * dd[..., 0] + dd[..., 1] * dd[..., 1]) + dd[..., 2] * dd[..., 2]).astype(np.float32)
            np.fill_diagonal(full, np.inf)
            near = np.sort(full, axis=1)[:, :k]
            want = (np.sqrt(near).astype(np.float64).cumsum(axis=1)[:, -1] / k).astype(np.float32)
            assert np.array_equal(md, want)

    check()
