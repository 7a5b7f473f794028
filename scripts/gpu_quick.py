"""Quick GPU shake-out: NN / ICP / outer-loop parity against the oracle + rough timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
from oracle import oracle_py as O

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    d = synth.make_pair(n)
    ctx = P.Context(0)
    t0 = time.time(); ctx.upload_pair(d); print("upload+build s", time.time() - t0)
    # NN parity
    q = np.concatenate([d["ct2"], d["bp2"]])
    idx, d2 = ctx.nn(q); print("nn kernel ms", ctx.last_device_ms(), "queries", len(q))
    oi, od = O.nn(d["ct1"], q)
    print("NN idx mismatches", int((idx != oi).sum()), "d2 mismatches", int((d2 != od).sum()))
    # far / outside queries
    rng = np.random.default_rng(1)
    qf = (rng.normal(0, 1, (5000, 3)) * [40, 40, 10]).astype(np.float32)
    idx, d2 = ctx.nn(qf); print("far nn ms", ctx.last_device_ms())
    oi, od = O.nn(d["ct1"], qf)
    print("far NN idx mismatches", int((idx != oi).sum()), "d2 mismatches", int((d2 != od).sum()))
    # full-cloud NN
    idx, d2 = ctx.nn(d["cloud2"], P.TGT_CLOUD1); print("cloud nn ms", ctx.last_device_ms(), len(d["cloud2"]))
    oi, od = O.nn(d["cloud1"], d["cloud2"])
    print("cloud NN idx mismatches", int((idx != oi).sum()), "d2 mismatches", int((d2 != od).sum()))
    # ICP parity (bit-exact against the oracle run with the same reduction geometry)
    ctx.icp_source_all()
    r = ctx.icp_run(P.icp_params(max_iter=20, force_iters=1), trace=True)
    print("icp ms", r["device_ms"], "iters", r["n_iter"], "geom", r["grid_blocks"], r["warps_per_block"],
          "Gcorr/s", r["correspondences"] / r["device_ms"] / 1e6)
    perm = ctx.icp_order()
    o = O.icp(d["ct1"], d["nrm1"], d["ct2"][perm], O.icp_params(max_iter=20, force_iters=1, reduce_mode=2,
              group_batches=r["group_batches"]), trace=True)
    print("ICP T bit-equal", np.array_equal(r["T"], o["T"]), "idx trace equal", np.array_equal(r["idx_trace"][:, perm], o["idx_trace"]),
          "T_trace equal", np.array_equal(r["T_trace"], o["T_trace"]), "mse equal", np.array_equal(r["mse"], o["mse"]))
    if not np.array_equal(r["T_trace"], o["T_trace"]):
        bad = [k for k in range(len(o["T_trace"])) if not np.array_equal(r["T_trace"][k], o["T_trace"][k])]
        print(" first differing iter", bad[:3], np.abs(r["T_trace"] - o["T_trace"]).max())
    o0 = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=20, force_iters=1), trace=False)
    print("vs sequential oracle: max|dT|", np.abs(r["T"] - o0["T"]).max())
    r2 = ctx.icp_run(); o2 = O.icp(d["ct1"], d["nrm1"], d["ct2"])
    print("default ICP iters gpu/oracle", r2["n_iter"], o2["n_iter"], "state", r2["state"], o2["state"], "max|dT|", np.abs(r2["T"] - o2["T"]).max())
    # outer loop
    pp = P.PairParams(d["Res1"], d["Res2"], d["SVRes1"], d["SVRes2"], d["DTmin"])
    t0 = time.time(); g = ctx.piecewise_icp(pp, 1, 0.05); tg = time.time() - t0
    pd = O.PairData(d)
    t0 = time.time(); o = O.piecewise_icp(pd, 1, 0.05); to = time.time() - t0
    print("outer gpu s", tg, "device ms", g["device_ms"], "oracle s", to)
    print("DTseries gpu", g["DTseries"]); print("DTseries orc", o["DTseries"])
    print("T max diff", np.abs(g["T"] - o["T"]).max(), "err vs truth", np.abs(g["T"] - d["T_true"]).max())
    print("VCM rel diff", np.abs(g["VCM"] - o["VCM"]).max() / np.abs(o["VCM"]).max())
    for a, b in zip(g["stats"], o["stats"]):
        print(" stable", a.n_stable, b.n_stable, "pts", a.n_stable_pts, b.n_stable_pts, "icp", a.icp_iters, b.icp_iters,
              "bb", a.maxBBchange, b.maxBBchange, "P75", a.P75, b.P75, "ms", a.device_ms)
    dl = ctx.source_download()
    print("cloud2 equal", np.array_equal(dl["cloud2"], pd.cloud2), "ct2 equal", np.array_equal(dl["ct2"], pd.ct2),
          "bp2 equal", np.array_equal(dl["bp2"], pd.bp2), "patch equal", np.array_equal(dl["patch_pts2"], pd.patch_pts2))
    print("launches", ctx.launch_count())

main()
