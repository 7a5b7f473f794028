"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libpwicp.so),
against the CPU oracle on the same seeded inputs, against the committed golden vectors, and --
at BASELINE.json's full sizes -- through size-independent properties.

Bars (BASELINE.json north_star): correspondence indices and float squared distances bit-exact;
rotation / translation within 1e-6 rad / 1e-6 m of the reference-order oracle; the whole inner
loop bit-exact against the oracle run in the device's summation order.
"""
import os

import time

import numpy as np
import pytest

import pwicp_b200 as P
from pwicp_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pair2k.npz")
POSE_TOL_RAD = 1e-6     # north_star: rotation within 1e-6 rad
POSE_TOL_M = 1e-6       # north_star: translation within 1e-6 m


def pose_diff(Ta, Tb):
    a, b = P.matrix2angle(Ta), P.matrix2angle(Tb)
    return float(np.abs(a - b).max()), float(np.abs(np.asarray(Ta)[:3, 3] - np.asarray(Tb)[:3, 3]).max())


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def pair2k():
    return synth.make_pair(2000, seed=20250606)


@pytest.fixture(scope="module")
def pair60k():
    return synth.make_pair(60000, seed=777)


# ------------------------------------------------------------------------------------------ A1
def test_nn_matches_oracle_bit_exact(gpu_ctx, oracle, pair60k):
    d = pair60k
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    q = np.concatenate([d["ct2"], d["bp2"]])
    idx, d2 = gpu_ctx.nn(q)
    oi, od = oracle.nn(d["ct1"], q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2, od)


def test_nn_matches_reference_kdtree(gpu_ctx, oracle, pair60k):
    """The device search against reference-authored code (oracle/_ref: the reference's codelibrary KD-tree)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_kdtree.so not built")
    d = pair60k
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    q = np.concatenate([d["ct2"], d["bp2"]])
    idx, d2 = gpu_ctx.nn(q)
    ri, rd = oracle.ref_nn(d["ct1"], q)
    bad = np.nonzero(ri != idx)[0]                      # only rounding-level ties may differ (double vs float metric)
    dr = q[bad] - d["ct1"][ri[bad]]
    fr = ((dr[:, 0] * dr[:, 0] + dr[:, 1] * dr[:, 1]) + dr[:, 2] * dr[:, 2]).astype(np.float32)
    assert (fr >= d2[bad]).all() and (fr <= d2[bad] * np.float32(1 + 3e-7)).all()          # within 2 ulp in float
    assert len(bad) < 5 and np.allclose(d2, rd, rtol=3e-7, atol=1e-12)


def test_nn_golden_vectors(gpu_ctx, gold, pair2k):
    d = pair2k
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    idx, d2 = gpu_ctx.nn(np.concatenate([d["ct2"], d["bp2"]]))
    assert np.array_equal(idx, gold["nn_pair_idx"]) and np.array_equal(d2, gold["nn_pair_d2"])
    rng = np.random.default_rng(7)
    tgt = rng.uniform(-5, 5, (100000, 3)).astype(np.float32)
    qry = rng.uniform(-6, 6, (20000, 3)).astype(np.float32)
    gpu_ctx.target_upload(tgt)
    idx, d2 = gpu_ctx.nn(qry)
    assert np.array_equal(idx, gold["nn_rand_idx"]) and np.array_equal(d2, gold["nn_rand_d2"])


@pytest.mark.parametrize("n1", [1, 2, 3, 5, 17, 300])
def test_nn_tiny_targets(gpu_ctx, oracle, n1):
    rng = np.random.default_rng(n1)
    tgt = rng.normal(0, 1, (n1, 3)).astype(np.float32)
    qry = rng.normal(0, 2, (257, 3)).astype(np.float32)
    gpu_ctx.target_upload(tgt)
    idx, d2 = gpu_ctx.nn(qry)
    oi, od = oracle.nn(tgt, qry, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_nn_edge_cases(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    # exact ties and duplicates -> lowest original index
    base = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float32)
    tgt = np.tile(base, (50, 1))
    gpu_ctx.target_upload(tgt)
    idx, d2 = gpu_ctx.nn(np.zeros((40, 3), np.float32))
    assert (idx == 0).all() and (d2 == 1.0).all()
    qry = (base[rng.integers(0, 6, 300)] * np.float32(0.75)).astype(np.float32)
    idx, d2 = gpu_ctx.nn(qry)
    oi, od = oracle.nn(tgt, qry, brute=True)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # queries exactly on targets, degenerate (planar / collinear) targets, far-away queries
    flat = rng.uniform(-3, 3, (5000, 3)).astype(np.float32); flat[:, 2] = 0.25
    line = np.zeros((2000, 3), np.float32); line[:, 0] = np.linspace(-4, 4, 2000, dtype=np.float32)
    for tgt in (flat, line):
        gpu_ctx.target_upload(tgt)
        qry = np.concatenate([tgt[::7], rng.normal(0, 3, (2000, 3)).astype(np.float32),
                              (rng.normal(0, 1, (500, 3)) * 500).astype(np.float32)])
        idx, d2 = gpu_ctx.nn(qry)
        oi, od = oracle.nn(tgt, qry)
        assert np.array_equal(idx, oi) and np.array_equal(d2, od)
        assert (d2[: len(tgt[::7])] == 0).all()
    # empty query batch
    idx, d2 = gpu_ctx.nn(np.zeros((0, 3), np.float32))
    assert len(idx) == 0


def test_nn_partial_overlap_far_queries(gpu_ctx, oracle, pair60k):
    """Queries far from a dense surface (non-overlapping scan parts) stay exact."""
    d = pair60k
    gpu_ctx.target_upload(d["ct1"])
    rng = np.random.default_rng(5)
    q = d["ct2"][::5].copy()
    q[:, 2] += rng.uniform(0.3, 3.0, len(q)).astype(np.float32)
    q[::3, 0] += np.float32(20.0)
    idx, d2 = gpu_ctx.nn(q)
    oi, od = oracle.nn(d["ct1"], q)
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)


def test_nonfinite_input_is_rejected(gpu_ctx):
    tgt = np.zeros((10, 3), np.float32); tgt[3, 1] = np.nan
    with pytest.raises(P.PwicpError) as e:
        gpu_ctx.target_upload(tgt)
    assert e.value.status == -3
    gpu_ctx.target_upload(np.eye(3, dtype=np.float32))
    q = np.zeros((5, 3), np.float32); q[0, 0] = np.inf
    with pytest.raises(P.PwicpError) as e:
        gpu_ctx.nn(q)
    assert e.value.status == -3


# ------------------------------------------------------------------------------------- A3 - A6
def test_inner_loop_bit_exact_in_device_order(gpu_ctx, oracle, pair60k):
    d = pair60k
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    gpu_ctx.icp_source_upload(d["ct2"])
    r = gpu_ctx.icp_run(P.icp_params(max_iter=12, force_iters=1), trace=True)
    perm = gpu_ctx.icp_order()
    assert sorted(perm.tolist()) == list(range(len(d["ct2"])))
    o = oracle.icp(d["ct1"], d["nrm1"], d["ct2"][perm],
                   oracle.icp_params(max_iter=12, force_iters=1, reduce_mode=2,
                                     group_batches=r["group_batches"]),
                   trace=True)
    assert r["n_iter"] == o["n_iter"] == 12
    assert np.array_equal(r["idx_trace"][:, perm], o["idx_trace"])    # indices of every inner iteration
    assert np.array_equal(r["T_trace"], o["T_trace"])                 # every incremental transform
    assert np.array_equal(r["mse"], o["mse"])
    assert np.array_equal(r["T"], o["T"])


def test_inner_loop_split_launch_small_set(gpu_ctx, oracle, pair60k, monkeypatch):
    """The two-launch form of the loop (stand-alone search of iteration 1; used from 500k points on) at a size the
    oracle finishes in seconds: bit-exact against the oracle and against the single launch."""
    d = pair60k
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    gpu_ctx.icp_source_upload(d["ct2"])
    one = gpu_ctx.icp_run(P.icp_params(max_iter=12, force_iters=1), trace=True)
    monkeypatch.setenv("PWICP_SPLIT_MIN_POINTS", "0")
    two = gpu_ctx.icp_run(P.icp_params(max_iter=12, force_iters=1), trace=True)
    assert two["research_ms"] > 0.0 and one["research_ms"] == 0.0
    for k in ("idx_trace", "T_trace", "mse", "T"):
        assert np.array_equal(one[k], two[k]), k
    perm = gpu_ctx.icp_order()
    o = oracle.icp(d["ct1"], d["nrm1"], d["ct2"][perm],
                   oracle.icp_params(max_iter=12, force_iters=1, reduce_mode=2, group_batches=two["group_batches"], threads=8), trace=True)
    assert np.array_equal(two["idx_trace"][:, perm], o["idx_trace"])
    assert np.array_equal(two["T_trace"], o["T_trace"]) and np.array_equal(two["mse"], o["mse"])
    nat = gpu_ctx.icp_run()                                            # criteria in charge: the loop may end in either launch
    monkeypatch.delenv("PWICP_SPLIT_MIN_POINTS")
    ref = gpu_ctx.icp_run()
    assert nat["n_iter"] == ref["n_iter"] and nat["state"] == ref["state"] and np.array_equal(nat["T"], ref["T"])
    monkeypatch.setenv("PWICP_SPLIT_MIN_POINTS", "0")
    m1 = gpu_ctx.icp_run(P.icp_params(max_iter=1))                     # one iteration: nothing to split
    assert m1["n_iter"] == 1 and m1["research_ms"] == 0.0


def test_inner_loop_source_outside_the_target_box(gpu_ctx, oracle, pair60k):
    """A source set far outside the target's bounding box: every point clamps into boundary cells of the target grid
    (one huge bin of the processing order, balls that span the whole grid).  Must stay fast and exact."""
    d = pair60k
    src = (d["ct2"][:20000] + np.array([0.0, 0.0, 60.0], np.float32)).astype(np.float32)
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    gpu_ctx.icp_source_upload(src)
    t0 = time.perf_counter()
    r = gpu_ctx.icp_run(P.icp_params(max_iter=3, force_iters=1), trace=True)
    assert time.perf_counter() - t0 < 20.0
    perm = gpu_ctx.icp_order()
    assert np.array_equal(np.sort(perm), np.arange(len(src)))
    o = oracle.icp(d["ct1"], d["nrm1"], src[perm],
                   oracle.icp_params(max_iter=3, force_iters=1, reduce_mode=2, group_batches=r["group_batches"], threads=8), trace=True)
    assert np.array_equal(r["idx_trace"][:, perm], o["idx_trace"])
    assert np.array_equal(r["T_trace"], o["T_trace"]) and np.array_equal(r["mse"], o["mse"])


def test_candidate_cache_adversarial_clouds(gpu_ctx, oracle):
    """The per-query candidate cache of the inner loop (exact by a ball-coverage certificate) on
    inputs that are not a sampled surface: a volume-filling random cloud, exact duplicates, and a
    lattice whose cell centres are equidistant from eight targets (exact float ties -> lowest
    index).  Every index of every iteration, every transform and every MSE must equal the oracle's."""
    rng = np.random.default_rng(7)
    vol = rng.uniform(-1.0, 1.0, (40000, 3)).astype(np.float32)
    dup = vol[rng.choice(len(vol), 3000, replace=False)]                       # exact duplicates
    g = np.arange(-8, 9, dtype=np.float32) * 0.125
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.float32(3.0)   # exact binary fractions
    tgt = np.concatenate([vol, dup, lat]).astype(np.float32)
    tgt = tgt[rng.permutation(len(tgt))]
    nrm = rng.normal(0, 1, tgt.shape); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm.astype(np.float32)
    centres = (lat[(np.abs(lat - 3.0) < 0.9).all(1)] + np.float32(0.0625)).astype(np.float32)   # 8-way ties
    near = (vol[:20000] + rng.normal(0, 2e-3, (20000, 3))).astype(np.float32)
    onto = dup[:2000].copy()                                                   # distance 0, duplicate indices
    src0 = np.concatenate([near, centres, onto]).astype(np.float32)
    T0 = synth.rigid_matrix(2e-4, -1e-4, 3e-4, 1e-4, -2e-4, 1e-4)
    src = (src0 @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32)
    gpu_ctx.target_upload(tgt, nrm, np.full(len(tgt), 1e-4, np.float32))
    gpu_ctx.icp_source_upload(src)
    iters = 16
    r = gpu_ctx.icp_run(P.icp_params(max_iter=iters, force_iters=1), trace=True)
    perm = gpu_ctx.icp_order()
    o = oracle.icp(tgt, nrm, src[perm], oracle.icp_params(max_iter=iters, force_iters=1, reduce_mode=2,
                                                         group_batches=r["group_batches"]), trace=True)
    assert r["n_iter"] == o["n_iter"] == iters
    assert np.array_equal(r["idx_trace"][:, perm], o["idx_trace"])
    assert np.array_equal(r["T_trace"], o["T_trace"]) and np.array_equal(r["mse"], o["mse"])
    # the same source untransformed: the lattice centres sit exactly on 8-way ties in iteration 0
    gpu_ctx.icp_source_upload(src0)
    r = gpu_ctx.icp_run(P.icp_params(max_iter=6, force_iters=1), trace=True)
    perm = gpu_ctx.icp_order()
    o = oracle.icp(tgt, nrm, src0[perm], oracle.icp_params(max_iter=6, force_iters=1, reduce_mode=2,
                                                          group_batches=r["group_batches"]), trace=True)
    assert np.array_equal(r["idx_trace"][:, perm], o["idx_trace"]) and np.array_equal(r["T_trace"], o["T_trace"])


def test_inner_loop_pose_vs_reference_order(gpu_ctx, oracle, pair60k, gold, pair2k):
    for d in (pair60k, pair2k):
        gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
        gpu_ctx.icp_source_upload(d["ct2"])
        for prm_g, prm_o in ((P.icp_params(), oracle.icp_params()),
                             (P.icp_params(max_iter=10, force_iters=1), oracle.icp_params(max_iter=10, force_iters=1))):
            r = gpu_ctx.icp_run(prm_g)
            o = oracle.icp(d["ct1"], d["nrm1"], d["ct2"], prm_o)      # sequential sums, what PCL does
            assert r["n_iter"] == o["n_iter"] and r["state"] == o["state"]
            da, dt = pose_diff(r["T"], o["T"])
            assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
    r = gpu_ctx.icp_run(P.icp_params(max_iter=10, force_iters=1), trace=True)
    da, dt = pose_diff(r["T"], gold["icp_T"])
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
    assert np.allclose(r["mse"], gold["icp_mse"], rtol=1e-9)
    r = gpu_ctx.icp_run()
    assert r["n_iter"] == int(gold["icp_default_iters"]) and r["state"] == int(gold["icp_default_state"])


def test_inner_loop_host_buffer_call_and_errors(gpu_ctx, oracle, pair2k):
    d = pair2k
    r = gpu_ctx.icp_p2plane(d["ct1"], d["nrm1"], d["ct2"])
    o = oracle.icp(d["ct1"], d["nrm1"], d["ct2"])
    da, dt = pose_diff(r["T"], o["T"])
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M and r["n_iter"] == o["n_iter"]
    with pytest.raises(P.PwicpError) as e:                            # < 3 correspondences
        gpu_ctx.icp_p2plane(d["ct1"], d["nrm1"], d["ct2"][:2])
    assert e.value.status == -6
    r1 = gpu_ctx.icp_p2plane(d["ct1"], d["nrm1"], d["ct2"][:3], P.icp_params(max_iter=1))
    assert r1["n_iter"] == 1 and r1["state"] == 1
    # the normals travel last and are checked on the device while the loop is already enqueued: the verdict must still
    # arrive with the result, for each of the three arrays, and the context must stay usable
    for key in ("nrm1", "ct2", "ct1"):
        bad = {k: d[k].copy() for k in ("ct1", "nrm1", "ct2")}
        bad[key][len(bad[key]) // 2, 1] = np.nan
        with pytest.raises(P.PwicpError) as e:
            gpu_ctx.icp_p2plane(bad["ct1"], bad["nrm1"], bad["ct2"])
        assert e.value.status == -3
    r2 = gpu_ctx.icp_p2plane(d["ct1"], d["nrm1"], d["ct2"])
    assert np.array_equal(r2["T"], r["T"]) and r2["n_iter"] == r["n_iter"]


# ------------------------------------------------------------------------------- A2, A7, A8, A10
def test_single_iteration_matches_oracle(gpu_ctx, oracle, pair60k):
    d = pair60k
    gpu_ctx.upload_pair(d)
    pd = oracle.PairData(d)
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    gs, os_ = P.State(0.05, 0, 0, 0, 0), oracle.State(0.05, 0, 0, 0, 0)
    for k in range(3):
        T, V, flags, st = gpu_ctx.single_iteration(pp, gs)
        rc, To, Vo, fo, so = oracle.single_iteration(pd, os_)
        assert rc == 0
        assert np.array_equal(flags, fo)                              # classification, patch by patch
        assert st.n_stable == so.n_stable and st.n_stable_pts == so.n_stable_pts
        assert st.LoDet_min == so.LoDet_min and st.LoDet_max == so.LoDet_max
        assert st.icp_iters == so.icp_iters and st.icp_state == so.icp_state
        da, dt = pose_diff(T, To)
        assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
        assert np.allclose(list(st.bb6), list(so.bb6), rtol=0, atol=1e-6)
        assert abs(st.maxBBchange - so.maxBBchange) <= 1e-6
        if not np.isnan(so.P75):
            assert abs(st.P75 - so.P75) <= 1e-6
        assert abs(gs.currDT - os_.currDT) <= 1e-7
        assert (gs.toStage2, gs.toStage3) == (os_.toStage2, os_.toStage3)
    dl = gpu_ctx.source_download()
    for k, ref in (("cloud2", pd.cloud2), ("ct2", pd.ct2), ("bp2", pd.bp2), ("patch_pts2", pd.patch_pts2)):
        assert np.abs(dl[k] - ref).max() <= 2e-6


def test_outer_loop_matches_oracle_and_golden(gpu_ctx, oracle, gold, pair2k):
    d = pair2k
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    for manual, dtinit, key in ((1, 0.05, "outer"), (0, 0.0, "outer_auto")):
        gpu_ctx.upload_pair(d)
        g = gpu_ctx.piecewise_icp(pp, manual, dtinit)
        o = oracle.piecewise_icp(oracle.PairData(d), manual, dtinit)
        assert g["n_outer"] == o["rc"]
        assert np.allclose(g["DTseries"], o["DTseries"], rtol=1e-6, atol=0)
        assert np.allclose(g["DTseries"], gold[key + "_DTseries"], rtol=1e-6, atol=0)
        assert (np.diff(g["DTseries"]) <= 0).all()
        da, dt = pose_diff(g["T"], o["T"])
        assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
        da, dt = pose_diff(g["T"], gold[key + "_T"])
        assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
        assert [s.n_stable for s in g["stats"]] == [s.n_stable for s in o["stats"]]
        assert [s.icp_iters for s in g["stats"]] == [s.icp_iters for s in o["stats"]]
        assert np.allclose(g["VCM"], o["VCM"], rtol=1e-6, atol=0)      # relative 1e-6
    assert np.allclose(g["VCM"], g["VCM"].T, rtol=1e-9)
    # a series against one reference epoch: the reference side stays resident, only the moving epoch is uploaded again
    # (pwicp_clouds_upload with cloud1 = NULL) -- same result, bit for bit
    gpu_ctx.upload_source_side(d)
    g2 = gpu_ctx.piecewise_icp(pp, 0, 0.0)
    assert np.array_equal(g2["T"], g["T"]) and np.array_equal(g2["DTseries"], g["DTseries"]) and np.array_equal(g2["VCM"], g["VCM"])
    fresh = P.Context(0)
    try:
        with pytest.raises(P.PwicpError) as e:
            fresh.clouds_upload(None, d["cloud2"])
        assert e.value.status == -2
    finally:
        fresh.close()


def test_outer_loop_error_codes(gpu_ctx, pair2k):
    d = dict(pair2k)
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    gpu_ctx.upload_pair(d)
    gpu_ctx.source_upload(d["ct2"][:3], d["bp2"][:18], d["bpstd2"][:3], d["patch_off2"][:4], d["patch_pts2"][: d["patch_off2"][3]])
    with pytest.raises(P.PwicpError) as e:
        gpu_ctx.single_iteration(pp, P.State(0.05, 0, 0, 0, 0))
    assert e.value.status == -4                                        # src/Registration.cpp:728-731
    gpu_ctx.source_upload(d["ct2"] + np.float32(5), d["bp2"] + np.float32(5), d["bpstd2"], d["patch_off2"], d["patch_pts2"])
    with pytest.raises(P.PwicpError) as e:
        gpu_ctx.single_iteration(pp, P.State(0.05, 0, 0, 0, 0))
    assert e.value.status == -5                                        # src/Registration.cpp:864-867


def test_standalone_pieces(gpu_ctx, oracle, gold, pair2k):
    d = pair2k
    # percentile (A7), bit-identical value
    assert gpu_ctx.percentile_nn(d["cloud1"], d["cloud2"], 0.75) == oracle.percentile_nn(d["cloud1"], d["cloud2"], 0.75) == gold["p75"][0]
    for pct in (0.0, 0.5, 0.999):
        assert gpu_ctx.percentile_nn(d["cloud1"], d["cloud2"], pct) == oracle.percentile_nn(d["cloud1"], d["cloud2"], pct)
    # overlap ratio (F1)
    _, d2 = oracle.nn(d["cloud1"], d["cloud2"])
    ratio = np.float32(np.float32((np.sqrt(d2) < np.float32(0.05)).sum()) / np.float32(len(d2)))
    assert gpu_ctx.overlap_ratio(d["cloud1"], d["cloud2"], 0.05) == ratio
    # transform (A5) bit-exact, odd sizes exercise the vector tail
    T = synth.rigid_matrix(0.01, -0.02, 0.03, 0.1, 0.2, -0.3).astype(np.float32)
    for n in (1, 2, 3, 4, 5, 1023, 4099):
        p = d["cloud2"][:n]
        assert np.array_equal(gpu_ctx.transform(p, T), oracle.transform(p, T))
    # octree bounding cube (A7)
    for res in (0.01, 0.3):
        assert np.array_equal(gpu_ctx.octree_bbox(d["cloud2"], res), oracle.octree_bbox(d["cloud2"], res))
    assert np.array_equal(gpu_ctx.octree_bbox(d["cloud2"], 0.01), gold["octree_bb"])
    # VCM (A8)
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    src = d["ct2"][~d["changed"]][:500]
    V, sing = gpu_ctx.vcm(src)
    Vo, so = oracle.vcm(d["ct1"], d["nrm1"], src)
    assert np.allclose(V, Vo, rtol=1e-6, atol=0) and sing == so
    assert np.allclose(V, gold["vcm"], rtol=1e-6, atol=0)


# --------------------------------------------------------------------- full size (BASELINE configs)
def test_full_size_properties_1m(gpu_ctx, oracle):
    """configs[1]: 1M-centroid pair.  Size-independent properties + an oracle-checked sample."""
    d = synth.make_pair(1_000_000, with_clouds=False)
    n = len(d["ct1"])
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    # (a) every target point is its own nearest neighbour at distance 0 (the generator is tie-free)
    idx, d2 = gpu_ctx.nn(d["ct1"])
    assert np.array_equal(idx, np.arange(n, dtype=np.int32)) and not d2.any()
    # (b) the reported distance is the reference float expression of the reported pair
    idx, d2 = gpu_ctx.nn(d["ct2"])
    df = d["ct2"] - d["ct1"][idx]
    assert np.array_equal(d2, ((df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]) + df[:, 2] * df[:, 2]).astype(np.float32))
    # (c) a random sample against brute force over all 1M targets
    rng = np.random.default_rng(1)
    pick = rng.choice(len(d["ct2"]), 300, replace=False)
    oi, od = oracle.nn(d["ct1"], d["ct2"][pick], brute=True)
    assert np.array_equal(idx[pick], oi) and np.array_equal(d2[pick], od)
    # (d) whole set against the oracle KD-tree
    oi, od = oracle.nn(d["ct1"], d["ct2"])
    assert np.array_equal(idx, oi) and np.array_equal(d2, od)
    # (d2) whole set against the reference's own KD-tree (oracle/_ref, codelibrary/util/tree/kd_tree.h compiled from
    # /root/reference; metric in double, so indices may differ only where the two float distances are within rounding)
    if oracle.ref_available():
        ri, rd = oracle.ref_nn(d["ct1"], d["ct2"])
        bad = np.nonzero(ri != idx)[0]
        dr = d["ct2"][bad] - d["ct1"][ri[bad]]
        fr = ((dr[:, 0] * dr[:, 0] + dr[:, 1] * dr[:, 1]) + dr[:, 2] * dr[:, 2]).astype(np.float32)
        assert (fr >= d2[bad]).all() and (fr <= d2[bad] * np.float32(1 + 3e-7)).all()      # within 2 ulp in float
        assert len(bad) < 10 and np.allclose(d2, rd, rtol=3e-7, atol=1e-12)
    # (e) eight inner iterations (search, cache build, cached iterations), bit-exact in device order,
    # pose within tolerance in reference order
    gpu_ctx.icp_source_upload(d["ct2"])
    r = gpu_ctx.icp_run(P.icp_params(max_iter=8, force_iters=1), trace=True)
    perm = gpu_ctx.icp_order()
    o1 = oracle.icp(d["ct1"], d["nrm1"], d["ct2"][perm],
                    oracle.icp_params(max_iter=8, force_iters=1, reduce_mode=2, group_batches=r["group_batches"]), trace=True)
    assert np.array_equal(r["T_trace"], o1["T_trace"]) and np.array_equal(r["idx_trace"][:, perm], o1["idx_trace"])
    assert np.array_equal(r["mse"], o1["mse"])
    o0 = oracle.icp(d["ct1"], d["nrm1"], d["ct2"], oracle.icp_params(max_iter=8, force_iters=1))
    da, dt = pose_diff(r["T"], o0["T"])
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
    # (f) 50 forced iterations are deterministic run to run
    a = gpu_ctx.icp_run(P.icp_params(max_iter=50, force_iters=1))
    b = gpu_ctx.icp_run(P.icp_params(max_iter=50, force_iters=1))
    assert np.array_equal(a["T"], b["T"]) and a["correspondences"] == 50 * len(d["ct2"])


def test_stress_10m_centroids(gpu_ctx, oracle):
    """configs[4]: 10M-centroid pair.  Properties that do not need a 10M x 10M oracle pass, plus six
    inner iterations (search, cache build, cached) bit-exact against the oracle in device order."""
    d = synth.make_pair(10_000_000, with_clouds=False)
    n = len(d["ct1"])
    gpu_ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    idx, d2 = gpu_ctx.nn(d["ct1"])
    assert np.array_equal(idx, np.arange(n, dtype=np.int32)) and not d2.any()
    idx, d2 = gpu_ctx.nn(d["ct2"])
    df = d["ct2"] - d["ct1"][idx]
    assert np.array_equal(d2, ((df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]) + df[:, 2] * df[:, 2]).astype(np.float32))
    rng = np.random.default_rng(2)
    pick = rng.choice(len(d["ct2"]), 100, replace=False)
    oi, od = oracle.nn(d["ct1"], d["ct2"][pick], brute=True)
    assert np.array_equal(idx[pick], oi) and np.array_equal(d2[pick], od)
    gpu_ctx.icp_source_upload(d["ct2"])
    r = gpu_ctx.icp_run(P.icp_params(max_iter=6, force_iters=1), trace=True)
    perm = gpu_ctx.icp_order()
    o = oracle.icp(d["ct1"], d["nrm1"], d["ct2"][perm],
                   oracle.icp_params(max_iter=6, force_iters=1, reduce_mode=2, group_batches=r["group_batches"]), trace=True)
    assert np.array_equal(r["T_trace"], o["T_trace"]) and np.array_equal(r["mse"], o["mse"])
    assert np.array_equal(r["idx_trace"][:, perm], o["idx_trace"])
    a = gpu_ctx.icp_run(P.icp_params(max_iter=30, force_iters=1))
    b = gpu_ctx.icp_run(P.icp_params(max_iter=30, force_iters=1))
    assert np.array_equal(a["T"], b["T"]) and a["correspondences"] == 30 * len(d["ct2"])
    print(f"10M: 30 iterations {a['device_ms']:.2f} ms, {a['correspondences'] / a['device_ms'] / 1e6:.2f} G corr/s")


# ------------------------------------------------------------------------------------- F3 (8f)
def test_patch_stats_match_oracle(gpu_ctx, oracle):
    """pwicp_patch_stats (one launch per cloud) against the oracle's calPatchCTandBP / calPatchNormal /
    calPatchSTD restatement: centroids and boundary points bit-exact (sequential float sums, strict
    comparisons), normals and sigmas within float-libm tolerance."""
    rng = np.random.default_rng(11)
    d = synth.make_pair(60000, pts_per_patch=16)
    xyz, off = d["patch_pts2"], d["patch_off2"]
    g = gpu_ctx.patch_stats(xyz, off)
    o = oracle.patch_stats(xyz, off)
    assert np.array_equal(g["ct"], o["ct"]) and np.array_equal(g["bp"], o["bp"])
    assert np.array_equal(g["nrm_ok"], o["nrm_ok"])
    assert np.abs(g["nrm"] - o["nrm"]).max() < 2e-5                 # atan2f/cosf/sinf: CUDA vs host libm
    assert np.allclose(g["bpstd"], o["bpstd"], rtol=1e-4) and np.allclose(g["ctstd"], o["ctstd"], rtol=1e-4)
    # ragged patch sizes, patches too small for a normal, a degenerate (collinear) patch
    sizes = rng.integers(1, 60, 3000)
    sizes[:5] = [1, 2, 4, 5, 6]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    pts = (rng.normal(0, 1, (off[-1], 3)) * [0.05, 0.05, 0.002] + rng.uniform(-5, 5, 3)).astype(np.float32)
    line = np.linspace(0, 1, sizes[10], dtype=np.float32)
    pts[off[10]:off[11]] = np.stack([line, 2 * line, -line], 1)
    g = gpu_ctx.patch_stats(pts, off)
    o = oracle.patch_stats(pts, off)
    assert np.array_equal(g["ct"], o["ct"]) and np.array_equal(g["bp"], o["bp"])
    assert np.array_equal(g["nrm_ok"][:3], [0, 0, 0]) and np.array_equal(g["nrm"][:3], np.tile([0, 0, 1], (3, 1)))
    big = sizes >= 8
    big[10] = False                                                  # the collinear patch has no defined normal
    assert np.array_equal(g["nrm_ok"][big], o["nrm_ok"][big])
    dn = np.minimum(np.abs(g["nrm"] - o["nrm"]).max(1), np.abs(g["nrm"] + o["nrm"]).max(1))
    assert dn[big].max() < 1e-3                                      # near-isotropic small patches amplify libm ulps
    ok = big & (sizes >= 3)
    assert np.allclose(g["bpstd"][ok], o["bpstd"][ok], rtol=1e-3)
    with pytest.raises(P.PwicpError):
        gpu_ctx.patch_stats(pts, off[::-1].copy())


def test_reference_pair_reproduces_recorded_result(gpu_ctx, oracle):
    """BASELINE configs[0] on the device: the reference's shipped Epoch_001 -> Epoch_002 pair at the hot-path boundary
    (tests/golden/refpair_e2.npz, segmented by the reference's own supervoxel code), per-patch constants from the device
    (pwicp_patch_stats), outer loop on the device (pwicp_piecewise_icp).  The result must be the 4x4 the reference recorded
    for this pair (results/4DPCReg/2_Direct2Ref_TransMatrix.txt) within 1e-6 rad / 1e-6 m, and the oracle's."""
    from conftest import load_refpair
    f = load_refpair(gpu_ctx.patch_stats)
    d = f["pair"]
    gpu_ctx.upload_pair(d)
    g = gpu_ctx.piecewise_icp(P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"]), 1, f["DTinit"])
    T = P.mat4_mul(P.mat4_mul(f["Sinv"], g["T"]), f["S"])                       # src/Registration.cpp:461
    da, dt = pose_diff(T, f["T_recorded"].astype(np.float32))
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M, (da, dt)
    sd, sd_rec = np.sqrt(np.diag(g["VCM"])), np.sqrt(np.diag(f["VCM_recorded"]))
    assert np.allclose(sd, sd_rec, rtol=2e-3)
    # the same centroid-level pair (device patch constants) through the oracle: identical schedule, pose within tolerance
    o = oracle.piecewise_icp(oracle.PairData(d), 1, f["DTinit"])
    assert np.array_equal(g["DTseries"], o["DTseries"])
    assert [s.n_stable for s in g["stats"]] == [s.n_stable for s in o["stats"]]
    da, dt = pose_diff(g["T"], o["T"])
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M, (da, dt)


# ---------------------------------------------------------------- F4: PCpreprocessing (VoxelGrid + StatisticalOutlierRemoval)
def _prep_clouds():
    rng = np.random.default_rng(42)
    scan = synth.make_scan(extent=2.0, spacing=0.005, seed=9)                      # a sampled surface, ~160k points
    vol = rng.uniform(-1, 1, (40000, 3)).astype(np.float32)                         # volume filling
    neg = (rng.normal(0, 0.3, (30000, 3)) + (-5.0, 7.0, -0.2)).astype(np.float32)   # negative / offset coordinates
    dup = np.repeat(rng.uniform(0, 1, (3000, 3)).astype(np.float32), 4, axis=0)     # exact duplicates
    return {"scan": scan, "volume": vol, "offset": neg, "duplicates": dup}


def test_voxel_grid_matches_oracle(gpu_ctx, oracle):
    """pcl::VoxelGrid (src/CommonFunc.cpp:430-433): same voxels, same order, bit-identical float centroids."""
    for name, c in _prep_clouds().items():
        for leaf in (0.005, 0.02, 0.3):
            g, o = gpu_ctx.voxel_grid(c, leaf), oracle.voxel_grid(c, leaf)
            assert g.shape == o.shape and np.array_equal(g, o), (name, leaf)
    one = np.array([[0.1, -0.2, 0.3]], np.float32)
    assert np.array_equal(gpu_ctx.voxel_grid(one, 0.01), one)
    same = np.tile(np.array([[0.101, 0.102, 0.103]], np.float32), (1000, 1)) + np.float32(1e-4) * np.arange(1000, dtype=np.float32)[:, None] / 1000
    g, o = gpu_ctx.voxel_grid(same, 1.0), oracle.voxel_grid(same, 1.0)              # everything in one voxel: 1000-term float sum
    assert len(g) == 1 and np.array_equal(g, o)
    with pytest.raises(P.PwicpError):
        gpu_ctx.voxel_grid(one, 0.0)


def test_knn_mean_dist_matches_oracle(gpu_ctx, oracle):
    """First pass of pcl::StatisticalOutlierRemoval (src/CommonFunc.cpp:446-451): bit-identical mean k-NN distances."""
    clouds = _prep_clouds()
    for name, c in clouds.items():
        for k in ((14, 1, 8, 20, 32) if name == "scan" else (14,)):
            g, o = gpu_ctx.knn_mean_dist(c, k), oracle.knn_mean_dist(c, k)
            assert np.array_equal(g, o), (name, k, int((g != o).sum()))
    tiny = clouds["volume"][:15]                                                    # n = k + 1: every other point is a neighbour
    assert np.array_equal(gpu_ctx.knn_mean_dist(tiny, 14), oracle.knn_mean_dist(tiny, 14))
    far = np.concatenate([clouds["volume"][:5000], np.array([[50, 50, 50], [-80, 3, 9]], np.float32)])   # isolated outliers
    assert np.array_equal(gpu_ctx.knn_mean_dist(far, 14), oracle.knn_mean_dist(far, 14))
    with pytest.raises(P.PwicpError):
        gpu_ctx.knn_mean_dist(tiny[:14], 14)


def test_preprocess_matches_oracle(gpu_ctx, oracle):
    """PCpreprocessing(cloud, out, true, Res, 14, 5.0) as the 4D driver calls it (src/Registration.cpp:415-416) and with the
    pairwise driver's multiplier 2.7 (:272-273); SOR alone (isDownSamp = false)."""
    rng = np.random.default_rng(5)
    scan = synth.make_scan(extent=2.0, spacing=0.005, seed=10)
    noisy = np.concatenate([scan, (scan[rng.choice(len(scan), 300)] + rng.normal(0, 0.05, (300, 3))).astype(np.float32)])
    for mult in (5.0, 2.7):
        g, o = gpu_ctx.preprocess(noisy, 0.005, 14, mult), oracle.preprocess(noisy, 0.005, 14, mult)
        assert g.shape == o.shape and np.array_equal(g, o) and len(g) < len(noisy)
    md = oracle.knn_mean_dist(noisy, 14)
    o, _ = oracle.sor_select(noisy, md, 5.0)
    assert np.array_equal(gpu_ctx.preprocess(noisy, 0.0, 14, 5.0, downsample=False), o)
    # full size: 1M points, bit-identical
    big = synth.make_pair(1_000_000, with_clouds=False)["ct1"]
    assert np.array_equal(gpu_ctx.knn_mean_dist(big, 14), oracle.knn_mean_dist(big, 14))
    assert np.array_equal(gpu_ctx.voxel_grid(big, 0.08), oracle.voxel_grid(big, 0.08))


# ------------------------------------------------------------------------- round 2: parity at size
def test_self_nn_matches_oracle_and_brute_force(gpu_ctx, oracle):
    """pwicp_self_nn (calPCresolution, src/CommonFunc.cpp:239-263: second neighbour of nearestKSearch(i, 2)): squared distance
    of every point to its nearest OTHER point.  Brute force in the reference's float expression at 4k points (duplicates
    included: distance 0 to the twin), the oracle's k-NN (k = 1: float(sqrt(double(d2)))) at 200k."""
    rng = np.random.default_rng(11)
    p = rng.uniform(-2, 2, (4000, 3)).astype(np.float32)
    p[100:140] = p[:40]                                                   # exact duplicates
    d2 = gpu_ctx.self_nn(p)
    dx = p[:, None, 0] - p[None, :, 0]; dy = p[:, None, 1] - p[None, :, 1]; dz = p[:, None, 2] - p[None, :, 2]
    bf = (dx * dx + dy * dy) + dz * dz                                   # float32, ((dx*dx)+dy*dy)+dz*dz
    np.fill_diagonal(bf, np.inf)
    assert np.array_equal(d2, bf.min(1))
    d = synth.make_pair(200_000, with_clouds=False)
    c = d["ct1"]
    d2 = gpu_ctx.self_nn(c)
    md = oracle.knn_mean_dist(c, 1)                                       # [PCL] float(dist_sum / k), dist = sqrt in double
    assert np.array_equal(np.sqrt(d2.astype(np.float64)).astype(np.float32), md)


def test_outer_loop_full_size_matches_oracle(gpu_ctx, oracle):
    """pwicp_piecewise_icp at BASELINE configs[1] size with clouds (1M patches, 8M patch points, 8M-point clouds) against the
    oracle (its independent NN queries on all host threads -- results do not depend on the thread count): identical DT series
    (the decreasing schedule of configs[4], all three stages), identical stable-set size and inner iteration count in every
    outer iteration, pose within 1e-6, VCM relative 1e-6."""
    d = synth.make_pair(1_000_000)
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    gpu_ctx.upload_pair(d)
    g = gpu_ctx.piecewise_icp(pp, 1, 0.05)
    threads = os.cpu_count() or 1
    oracle.set_threads(threads)
    try:
        o = oracle.piecewise_icp(oracle.PairData(d), 1, 0.05, oracle.icp_params(threads=threads))
    finally:
        oracle.set_threads(1)
    assert g["n_outer"] == o["rc"] >= 5
    assert np.array_equal(g["DTseries"], o["DTseries"])
    assert (np.diff(g["DTseries"]) <= 0).all() and g["DTseries"][-1] < g["DTseries"][0]
    assert [s.n_stable for s in g["stats"]] == [s.n_stable for s in o["stats"]]
    assert [s.n_stable_pts for s in g["stats"]] == [s.n_stable_pts for s in o["stats"]]
    assert [s.icp_iters for s in g["stats"]] == [s.icp_iters for s in o["stats"]]
    p75g = [s.P75 for s in g["stats"]]; p75o = [s.P75 for s in o["stats"]]
    assert [np.isnan(v) for v in p75g] == [np.isnan(v) for v in p75o]     # stage 1 ran in the same iterations ...
    assert np.array_equal(np.nan_to_num(p75g), np.nan_to_num(p75o))       # ... and its percentile is the same number
    assert any(np.isnan(v) for v in p75g) and not all(np.isnan(v) for v in p75g)     # stages 1 and 2 both ran
    assert g["stats"][-1].vcm_written == 1                                            # stage 3 reached
    da, dt = pose_diff(g["T"], o["T"])
    assert da <= POSE_TOL_RAD and dt <= POSE_TOL_M
    assert np.allclose(g["VCM"], o["VCM"], rtol=1e-6, atol=0)


def test_max_outer_is_reported(gpu_ctx, pair2k):
    d = pair2k
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    gpu_ctx.upload_pair(d)
    with pytest.raises(P.PwicpError) as e:
        gpu_ctx.piecewise_icp(pp, 1, 0.05, max_outer=1)
    assert e.value.status == -8                                           # PWICP_ERR_MAX_OUTER: no VCM, transformation so far


def test_knn_normals_match_oracle(gpu_ctx, oracle):
    """pwicp_knn_normals (front end of the segmentation, src/Segmentation.cpp:28-46: kNN = 45 neighbours + PCAEstimateNormal per
    point) against the oracle's restatement (itself bit-identical to the reference's codelibrary, tests/test_oracle.py): the
    neighbour lists index for index (double metric, ties by index, exact duplicates included), the normals within libm
    tolerance (pow / acos / cos of the closed-form eigenvalue are not correctly rounded on either side)."""
    pts = synth.make_scan(extent=2.0, spacing=0.01, seed=5)                  # 40k points of a dense scan
    pts[1000:1040] = pts[:40]                                                 # exact duplicates: distance ties at 0
    nb, nr = gpu_ctx.knn_normals(pts, 45)
    onb, onr = oracle.knn_normals(pts, 45)
    assert np.array_equal(nb, onb)
    assert np.abs(nr - onr).max() <= 1e-9
    assert np.allclose(np.linalg.norm(nr, axis=1), 1.0, atol=1e-12)
    nb8, _ = gpu_ctx.knn_normals(pts[:5000], 8)
    onb8, _ = oracle.knn_normals(pts[:5000], 8)
    assert np.array_equal(nb8, onb8)
