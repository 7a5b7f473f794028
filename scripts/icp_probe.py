"""Inner-loop probe: bit-exact parity against the oracle at a small size, then per-iteration device times and the
number of queries that ran the search at the bench size (pwicp_icp_profile).  python scripts/icp_probe.py [n] [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))
import numpy as np
import pwicp_b200 as P
from pwicp_b200 import synth
from oracle import oracle_py as O


def parity(ctx, n, iters):
    d = synth.make_pair(n, with_clouds=False)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])
    r = ctx.icp_run(P.icp_params(max_iter=iters, force_iters=1), trace=True)
    perm = ctx.icp_order()
    o = O.icp(d["ct1"], d["nrm1"], d["ct2"][perm],
              O.icp_params(max_iter=iters, force_iters=1, reduce_mode=2, group_batches=r["group_batches"], threads=16), trace=True)
    ok_idx = np.array_equal(r["idx_trace"][:, perm], o["idx_trace"])
    ok_T = np.array_equal(r["T_trace"], o["T_trace"])
    ok_mse = np.array_equal(r["mse"], o["mse"])
    print(f"parity n={len(d['ct2'])} iters={iters} grid={r['grid_blocks']}x{r['warps_per_block']}: idx {ok_idx} T {ok_T} mse {ok_mse}"
          f" natural_iters={r['natural_iters']} state={r['natural_state']}", flush=True)
    if not ok_T:
        bad = [k for k in range(len(o["T_trace"])) if not np.array_equal(r["T_trace"][k], o["T_trace"][k])]
        print("  first differing iterations", bad[:5], "max|dT|", np.abs(r["T_trace"] - o["T_trace"]).max())
    if not ok_idx:
        print("  idx mismatches per iteration", (r["idx_trace"][:, perm] != o["idx_trace"]).sum(1))
    return ok_idx and ok_T and ok_mse


def timing(ctx, n, iters):
    d = synth.make_pair(n, with_clouds=False)
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])
    prm = P.icp_params(max_iter=iters, force_iters=1)
    for _ in range(2):
        ctx.flush_l2(); ctx.target_rebuild(); r = ctx.icp_run(prm)
    best = None
    for _ in range(5):
        ctx.flush_l2(); b = ctx.target_rebuild(); r = ctx.icp_run(prm)
        if best is None or r["device_ms"] < best[1]["device_ms"]:
            best = (b, r, ctx.icp_profile(), ctx.icp_phase_profile())
    b, r, (us, srch), ph = best
    n2 = len(d["ct2"])
    print(f"timing n={n2} iters={iters}: build {b:.3f} ms, loop {r['device_ms']:.3f} ms, kernel {r['kernel_ms']:.3f} ms + iteration-1 search {r['research_ms']:.3f} ms "
          f"-> {r['correspondences'] / (b + r['device_ms']) / 1e6:.2f} G corr/s resident; kernel alg. "
          f"{48 * iters * n2 / r['kernel_ms'] / 1e6:.0f} GB/s", flush=True)
    print("  iteration us :", " ".join(f"{u:.1f}" for u in us[:8]), "... median of the rest", f"{np.median(us[8:]):.2f}" if len(us) > 8 else "")
    if len(ph) > 10:
        print("  CTA 0 phases (median of iterations 8..): own batches done %.2f us, CTA sum posted %.2f, totals %.2f, solved %.2f"
              % tuple(np.median(ph[8:], axis=0)))
    print("  searched     :", " ".join(str(s) for s in srch[:8]), "... sum of the rest", int(srch[8:].sum()))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    if os.environ.get("PWICP_LIB"):
        P._lib = P.load_library(os.environ["PWICP_LIB"]); print("library", os.environ["PWICP_LIB"])
    ctx = P.Context(0)
    ok = parity(ctx, 60_000, 12)
    ok = parity(ctx, 2_000, 8) and ok
    timing(ctx, n, iters)
    if len(sys.argv) > 3:
        timing(ctx, int(sys.argv[3]), iters)
    print("PROBE", "OK" if ok else "FAILED")
