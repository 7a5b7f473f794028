"""Generates the golden fixtures of tests/golden/ from the CPU oracle (the reference
holds no fixture at the hot-path boundary, so the build pins its own).

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by pwicp_b200.synth, only OUTPUTS are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))

from oracle import oracle_py as O          # noqa: E402
from pwicp_b200 import synth               # noqa: E402

N_PAIR = 2000
SEED = 20250606


def random_pair(n1=100000, nq=20000, seed=7):
    rng = np.random.default_rng(seed)
    tgt = rng.uniform(-5, 5, (n1, 3)).astype(np.float32)
    qry = rng.uniform(-6, 6, (nq, 3)).astype(np.float32)
    return tgt, qry


def main():
    d = synth.make_pair(N_PAIR, seed=SEED)
    out = {}
    # (i) NN vectors: centroid pair + a 100k random pair
    q = np.concatenate([d["ct2"], d["bp2"]])
    out["nn_pair_idx"], out["nn_pair_d2"] = O.nn(d["ct1"], q)
    tgt, qry = random_pair()
    out["nn_rand_idx"], out["nn_rand_d2"] = O.nn(tgt, qry)
    # (ii) ten forced inner iterations, sequential (reference) summation order
    r = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=10, force_iters=1), trace=True)
    out["icp_T"], out["icp_T_trace"], out["icp_mse"] = r["T"], r["T_trace"], r["mse"]
    out["icp_idx_hash"] = np.array([int(np.bitwise_xor.reduce(r["idx_trace"][k].astype(np.int64) * (np.arange(r["idx_trace"].shape[1]) + 1)))
                                    for k in range(r["n_iter"])], np.int64)
    ATA, ATb, x, T = O.lls_step(d["ct2"], O.nn(d["ct1"], d["ct2"])[0], d["ct1"], d["nrm1"])
    out["lls_ATA"], out["lls_ATb"], out["lls_x"], out["lls_T"] = ATA, ATb, x, T
    # default settings (convergence criteria active)
    r = O.icp(d["ct1"], d["nrm1"], d["ct2"])
    out["icp_default_T"], out["icp_default_iters"], out["icp_default_state"] = r["T"], r["n_iter"], r["state"]
    # (iii) full outer loop
    pd = O.PairData(d)
    res = O.piecewise_icp(pd, 1, 0.05)
    out["outer_DTseries"], out["outer_T"], out["outer_VCM"] = res["DTseries"], res["T"], res["VCM"]
    out["outer_n_stable"] = np.array([s.n_stable for s in res["stats"]], np.int32)
    out["outer_icp_iters"] = np.array([s.icp_iters for s in res["stats"]], np.int32)
    out["outer_P75"] = np.array([s.P75 for s in res["stats"]])
    out["outer_bb"] = np.array([s.maxBBchange for s in res["stats"]], np.float32)
    out["outer_ct2_final"] = pd.ct2[:50].copy()
    # automatic DTinit = 3 * P75(cloud1, cloud2)
    res2 = O.piecewise_icp(O.PairData(d), 0, 0.0)
    out["outer_auto_DTseries"], out["outer_auto_T"] = res2["DTseries"], res2["T"]
    # (iv) VCM + helpers
    out["vcm"], _ = O.vcm(d["ct1"], d["nrm1"], d["ct2"][~d["changed"]][:500])
    out["p75"] = np.array([O.percentile_nn(d["cloud1"], d["cloud2"], 0.75)])
    out["octree_bb"] = O.octree_bbox(d["cloud2"], 0.01)
    np.savez_compressed(os.path.join(HERE, "pair2k.npz"), **out)
    print("wrote pair2k.npz:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
