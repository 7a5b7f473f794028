#!/usr/bin/env python
"""bench.py -- headline benchmark of the Piecewise-ICP inner registration loop on B200.

Workload (BASELINE.json configs[1]): pairwise, 1M-centroid synthetic planar-patch pair, 50 inner point-to-plane ICP
iterations; the convergence criteria are evaluated every iteration but may not stop the loop.
One "step" = device build of the target grid + the 50-iteration inner loop over all source centroids (per-pair grid
build included, one-time uploads excluded).

  value  : correspondences/s with inputs already resident in HBM (CUDA events on the library stream)
  e2e    : the same metric through the host-buffer C-ABI call pwicp_icp_p2plane (the call shape of
           P2PICPwithPatchNormal), H2D of both clouds from pinned memory and D2H of the 4x4 inside the timed region
  phases : where a resident step goes (grid build / processing order / iteration-0 search pre-pass / stand-alone search of
           iteration 1 / persistent kernel: search iterations, cached iterations), `natural`: the same pair with the convergence criteria allowed to stop the loop
  N > 1  : the headline repeats per rank (one pair per rank, no data-path collective: weak scaling); the sharded
           workload that can fail to scale is `config4` below
  config4: BASELINE configs[3], "4D synthetic: 64 epochs x 2M pts, epoch-sharded": every epoch is registered against the
           reference epoch through the product's outer loop (the reference epoch resident, upload of the moving epoch from pinned
           host memory + pwicp_piecewise_icp),
           epochs dealt out as PiecewiseICP_4D_shard does ((step - 1) % world == rank), the fixed 384-byte records
           all-gathered over NCCL -- all inside one barrier-bracketed timed region (strong scaling; runs at every N, 1 included)

`--impl reference` times the CPU oracle (oracle/, the restatement of the reference's PCL path; the reference itself cannot be
built here) on all host threads, the same 50 iterations on the same pair per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "piecewise-icp_b200", "python"))

import numpy as np

N_CENTROIDS = 1_000_000
INNER_ITERS = 50
ALG_BYTES_PER_CORR = 48          # 12 src + 12 tgt xyz + 12 tgt normal + 12 write
STREAMED_BYTES_PER_CORR = 64     # what a cached iteration moves: three 16-byte reads + one 16-byte write
METRIC = "correspondences/s/GPU (ICP iters/s on 1M-pt pair; pose err vs ref)"
C4_EPOCHS = 64
C4_PATCHES = 250_000             # x 8 points per patch = 2M points per epoch


def workload_config(n1, n2, seed):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": "pairwise, 1M-point synthetic planar-patch cloud, 50 ICP iterations (BASELINE configs[1])",
            "n_target": int(n1), "n_source": int(n2), "inner_iters": INNER_ITERS,
            "convergence": "DefaultConvergenceCriteria evaluated every iteration, not allowed to stop the loop",
            "step": "target search structure built + 50 inner iterations over all source centroids",
            "l2": "flushed (384 MiB fill) between timed steps on the GPU arm; the pair (60 MB) exceeds no cache on the CPU arm",
            "seed": int(seed)}


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    same configuration (profiles/traffic.json, written by scripts/summarize_profile.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"]), t.get("source")
    except Exception:
        return None, None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons; `stop(t0, t1)` keeps the samples taken inside the
    timed region [t0, t1] (wall clock).  Started before the warm-up: nvidia-smi needs ~0.1 s to come up."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.rows
        window = "timed region"
        if t0 is not None:
            inside = [r for r in rows if t0 <= r[0] <= t1 + 0.03]
            if len(inside) >= 2:
                rows = inside
            else:                       # a very short timed region: the samples closest to it
                rows = sorted(rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:5]
                window = "nearest samples (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for _, r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 10:
                continue
            try:
                sm.append(float(f[2])); mx.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """CPU arm: the oracle's inner loop on the host cores, the whole workload of a GPU step per step: search structure
    (KD-tree, serial) built + 50 iterations on the full 1M x 1M pair.  The reference itself is single-threaded (no
    OpenMP / threads anywhere in its tree); here the independent NN queries and row terms use every host thread (sums
    stay sequential) and the single-thread figure is stated next to it."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    from pwicp_b200 import synth
    d = synth.make_pair(N_CENTROIDS, with_clouds=False)
    n1, n2 = len(d["ct1"]), len(d["ct2"])
    threads = os.cpu_count() or 1
    prm = O.icp_params(max_iter=INNER_ITERS, force_iters=1, threads=threads)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.icp(d["ct1"], d["nrm1"], d["ct2"], prm)        # tree build + 50 iterations
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = len(times) * INNER_ITERS * n2 / tot
    t0 = time.perf_counter()
    sample_iters = 10
    O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=sample_iters, force_iters=1))
    single = sample_iters * n2 / (time.perf_counter() - t0)
    sample = (f"all {INNER_ITERS} inner iterations on the full {n1}x{n2} pair, KD-tree build (serial) included, per step; "
              f"NN queries and row terms on {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "correspondences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
        "config": workload_config(n1, n2, synth.SEED_TARGET),
        "cpu_baseline": {"value": value, "unit": "correspondences/s", "cores": threads, "kind": "port",
                         "sample": sample, "single_thread_value": single,
                         "single_thread_sample": f"{sample_iters} iterations incl. the tree build, one thread (what the reference does)",
                         "note": "the reference is single-threaded; kind=port because PCL/Eigen/Boost are absent "
                                 "(DESIGN.md section 7)"},
        "e2e": {"value": value, "unit": "correspondences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config4_leg(P, synth, torch, dist, local_rank, rank, world):
    """BASELINE configs[3]: 64 epochs x 2M points, every epoch against the reference epoch, epoch-sharded."""
    base = synth.make_pair(C4_PATCHES, seed=synth.SEED_TARGET + 7)
    n2 = len(base["ct2"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    tgt = {k: pin(base[k]) for k in ("ct1", "nrm1", "ctstd1", "cloud1")}
    fixed = {k: pin(base[k]) for k in ("bpstd2", "patch_off2")}
    mine = [e for e in range(C4_EPOCHS) if e % world == rank]      # (step - 1) % world == rank: PiecewiseICP_4D_shard
    rng = np.random.default_rng(synth.SEED_SOURCE + 11)
    motions = [np.concatenate([rng.uniform(-0.004, 0.004, 3), rng.uniform(-0.01, 0.01, 3)]) for _ in range(C4_EPOCHS)]
    epochs = {}
    for e in mine:                                   # synthesis is not part of the path: before the timed region
        T = synth.rigid_matrix(*motions[e])
        mv = lambda a: pin((a.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
        epochs[e] = {"ct2": mv(base["ct2"]), "bp2": mv(base["bp2"]), "patch_pts2": mv(base["patch_pts2"])}
        epochs[e]["cloud2"] = epochs[e]["patch_pts2"]
    # two contexts per rank: while one registers epoch e, a loader thread uploads epoch e + 1 into the other (its three
    # grid builds run on that context's stream).  ctypes drops the GIL inside the library calls.
    ctxs = [P.Context(local_rank), P.Context(local_rank)]
    pp = P.PairParams(base["Res1"], base["Res2"], base["SVRes1"], base["SVRes2"], base["DTmin"])
    t_up, t_icp = [0.0], [0.0]

    have_ref = [False, False]

    def upload(k, e):
        # the reference epoch goes up once per context (inside the timed region), every other upload is the moving epoch
        d = dict(tgt); d.update(fixed); d.update(epochs[e])
        t0 = time.perf_counter()
        if have_ref[k]:
            ctxs[k].upload_source_side(d)
        else:
            ctxs[k].upload_pair(d); have_ref[k] = True
        t_up[0] += time.perf_counter() - t0

    def register(k):
        t0 = time.perf_counter()
        g = ctxs[k].piecewise_icp(pp, 1, 0.05)
        t_icp[0] += time.perf_counter() - t0
        return g

    for k in (0, 1):                                 # warm-up of both contexts: allocations, first launches
        upload(k, mine[0]); register(k)
    rec = torch.zeros((C4_EPOCHS, 96), dtype=torch.float32, device="cuda")       # 384-byte record per epoch
    if dist:                                         # ... and the first all-gather of this shape (NCCL sets up lazily)
        allrec = torch.zeros((world,) + tuple(rec.shape), dtype=rec.dtype, device="cuda")
        for _ in range(2):                           # every kernel of the tail once before the clock starts: CUDA loads
            dist.all_gather_into_tensor(allrec, rec)  # a module at its first launch (the reduction over the ranks cost
            _ = allrec.sum(0)                         # 9 ms in the timed region of r02x at 8 GPUs)
            rec[0] = torch.from_numpy(np.zeros(96, np.float32)).cuda()
            torch.cuda.synchronize()
        dist.barrier()
    t_up[0] = t_icp[0] = 0.0
    have_ref[0] = have_ref[1] = False                # the warm-up's reference uploads do not count
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev_ms, corr, outer = 0.0, 0, 0
    upload(0, mine[0])
    for j, e in enumerate(mine):
        loader = None
        if j + 1 < len(mine):
            loader = threading.Thread(target=upload, args=((j + 1) % 2, mine[j + 1]))
            loader.start()
        g = register(j % 2)
        dev_ms += g["device_ms"]; outer += int(g["n_outer"])
        corr += sum(int(s.icp_iters) * int(s.n_stable) for s in g["stats"])
        r = np.zeros(96, np.float32)
        r[:16] = g["T"].reshape(16); r[16:52] = g["VCM"].reshape(36).astype(np.float32); r[52] = 1.0
        rec[e] = torch.from_numpy(r).cuda()
        if loader:
            loader.join()
    torch.cuda.synchronize()
    t_loop = time.perf_counter() - t0
    if dist:
        dist.all_gather_into_tensor(allrec, rec)     # the one collective of the 4D mode: 384 bytes per epoch and rank
        rec = allrec.sum(0)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    wall_own = wall
    t_loop_min, t_after_loop_max = t_loop, wall - t_loop
    if dist:
        dist.barrier()
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX); wall = float(t[0])
        c = torch.tensor([dev_ms, corr, outer], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM); dev_ms, corr, outer = float(c[0]), float(c[1]), int(c[2])
        m = torch.tensor([t_up[0], t_icp[0], t_loop, -t_loop, wall_own - t_loop], dtype=torch.float64, device="cuda")
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        t_up[0], t_icp[0], t_loop, t_loop_min, t_after_loop_max = float(m[0]), float(m[1]), float(m[2]), -float(m[3]), float(m[4])
    done = int((rec[:, 52] > 0).sum().item())
    h2d_ref = int(sum(v.nbytes for v in tgt.values()))
    h2d_epoch = int(sum(v.nbytes for v in fixed.values()) +
                    sum(epochs[mine[0]][k].nbytes for k in ("ct2", "bp2", "patch_pts2", "cloud2")))
    # ground truth: epoch e was moved by motions[e]; the estimate maps it back
    e0 = mine[0]
    T_est = rec[e0, :16].cpu().numpy().reshape(4, 4).astype(np.float64)
    resid = T_est @ synth.rigid_matrix(*motions[e0]) @ synth.rigid_matrix(*synth.DEFAULT_MOTION)   # estimate o applied motion
    for c_ in ctxs:
        c_.close()
    return {"workload": "4D synthetic: 64 epochs x 2M pts each, every epoch against the reference epoch, epoch-sharded "
                        "(BASELINE configs[3])",
            "epochs": C4_EPOCHS, "patches_per_epoch": int(n2), "points_per_epoch": int(len(base["patch_pts2"])),
            "scaling": "strong", "n_gpus": world, "epochs_registered": done,
            "wall_s": wall, "epochs_per_s": C4_EPOCHS / wall, "device_ms_sum_over_ranks": dev_ms,
            "outer_iterations": outer, "correspondences": corr, "correspondences_per_s": corr / wall,
            "h2d_bytes_per_epoch": h2d_epoch, "h2d_bytes_reference_epoch_once_per_context": h2d_ref, "record_bytes_gathered": C4_EPOCHS * 384 * world,
            "slowest_rank_s": {"uploads_and_grid_builds": t_up[0], "outer_loops": t_icp[0], "epoch_loop": t_loop,
                               "epoch_loop_fastest_rank": t_loop_min,
                               "after_the_loop_max_over_ranks": t_after_loop_max},
            "upload_gbs_per_gpu_incl_grid_builds": (h2d_epoch * len(mine) + h2d_ref * min(2, len(mine))) / max(t_up[0], 1e-9) / 1e9,
            "timed_region": "barrier | the reference epoch once per context (pwicp_target_upload + cloud1: two device grid "
                            "builds) | per epoch: upload of the moving epoch from pinned host memory (pwicp_source_upload, "
                            "pwicp_clouds_upload with the resident cloud1) + pwicp_piecewise_icp, the "
                            "upload of the next epoch overlapping the registration of the current one (two contexts per rank, "
                            "a loader thread) | NCCL all-gather of the records | sync; wall clock, max over ranks",
            "sharding": "epoch e on rank e % world, as PiecewiseICP_4D_shard ((step - 1) % world == rank)",
            "residual_of_first_epoch_vs_truth": float(np.abs(resid - np.eye(4)).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--centroids", dest="n", type=int, default=N_CENTROIDS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import pwicp_b200 as P
    from pwicp_b200 import synth

    if not torch.cuda.is_available() or P.load_library().pwicp_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # each rank registers its own pair ("epoch"): same generator, rank-dependent seed
    d = synth.make_pair(args.n, seed=synth.SEED_TARGET + 100 * rank, with_clouds=False)
    n1, n2 = len(d["ct1"]), len(d["ct2"])
    ctx = P.Context(local_rank)
    prm = P.icp_params(max_iter=INNER_ITERS, force_iters=1)

    # ---- resident arm ---------------------------------------------------------------------
    ctx.target_upload(d["ct1"], d["nrm1"], d["ctstd1"])
    ctx.icp_source_upload(d["ct2"])

    def step_resident():
        ctx.flush_l2()                       # L2 hygiene between timed steps (not timed)
        b_ms = ctx.target_rebuild()
        r = ctx.icp_run(prm)
        return b_ms, r

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    wall0 = time.perf_counter()
    epoch0 = time.time()
    build_ms, icp_ms, kern_ms, sort_ms, pre_ms, res_ms, corr = [], [], [], [], [], [], 0
    it_search, it_cached, n_search = [], [], []
    last = None
    for _ in range(args.steps):
        b_ms, r = step_resident()
        build_ms.append(b_ms); icp_ms.append(r["device_ms"]); kern_ms.append(r["kernel_ms"]); corr += r["correspondences"]
        sort_ms.append(r["sort_ms"]); pre_ms.append(r["prepass_ms"]); res_ms.append(r["research_ms"])
        us, srch = ctx.icp_profile()
        searching = srch > 0.01 * n2                       # iterations in which more than 1 % of the queries ran the search
        it_search.append(float(us[searching].sum()) * 1e-3); n_search.append(int(searching.sum()))
        it_cached.append(float(np.median(us[~searching])) if (~searching).any() else float("nan"))
        last = r
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0 - args.steps    # minus the L2-flush fills
    clocks = sampler.stop(epoch0, time.time())
    dev_s = (sum(build_ms) + sum(icp_ms)) / 1e3

    # the same pair with the convergence criteria in charge (what a real call does)
    ctx.flush_l2(); nb_ms = ctx.target_rebuild()
    nat = ctx.icp_run(P.icp_params(max_iter=INNER_ITERS))
    natural = {"iterations": int(nat["n_iter"]), "state": P.CONV_NAMES.get(int(nat["state"]), str(nat["state"])),
               "device_ms": float(nb_ms + nat["device_ms"]),
               "correspondences_per_s": float(nat["correspondences"] / ((nb_ms + nat["device_ms"]) * 1e-3)),
               "forced_run_met_criteria_at_iteration": int(last["natural_iters"])}

    # ---- e2e arm: host buffers through the reference-shaped call ---------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_t, h_n, h_s = pin(d["ct1"]), pin(d["nrm1"]), pin(d["ct2"])
    for _ in range(2):
        ctx.icp_p2plane(h_t, h_n, h_s, prm)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_t = []
    for _ in range(max(10, args.steps // 2)):
        ctx.flush_l2()
        t0 = time.perf_counter()
        r2 = ctx.icp_p2plane(h_t, h_n, h_s, prm)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = sum(e2e_t)
    e2e_corr = len(e2e_t) * INNER_ITERS * n2
    # what the host link of this box delivers for the same pinned buffers (explains e2e - resident; boxes differ)
    dst = torch.empty(h_t.size, dtype=torch.float32, device="cuda")
    srcs = [torch.from_numpy(x).reshape(-1) for x in (h_t, h_n, h_s)]
    e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dst.copy_(srcs[0], non_blocking=True); torch.cuda.synchronize()
    e_a.record()
    for x in srcs:
        dst.copy_(x, non_blocking=True)
    e_b.record(); torch.cuda.synchronize()
    h2d_gbs = (h_t.nbytes + h_n.nbytes + h_s.nbytes) / (e_a.elapsed_time(e_b) * 1e-3) / 1e9
    h2d = int(h_t.nbytes + h_n.nbytes + h_s.nbytes)

    # ---- aggregate over ranks (max time, summed work) + the 4D-style record gather -----------
    rec = torch.zeros(96, dtype=torch.float32, device="cuda")   # 384-byte per-pair record
    rec[:16] = torch.from_numpy(last["T"].reshape(16)).cuda()
    if dist:
        t = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([corr, e2e_corr, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        corr, e2e_corr, launches = float(c[0]), float(c[1]), int(c[2])
        recs = [torch.zeros_like(rec) for _ in range(world)]
        dist.all_gather(recs, rec)
    value = corr / dev_s
    e2e_value = e2e_corr / e2e_s

    line = None
    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        traffic, traffic_src = profiled_traffic()
        icp_avg_ms = float(np.mean(icp_ms))
        kern_avg_ms = float(np.mean(kern_ms))
        achieved = ALG_BYTES_PER_CORR * INNER_ITERS * n2 / (kern_avg_ms * 1e-3) / 1e9
        cached_us = float(np.nanmedian(it_cached))
        line = {
            "metric": METRIC, "value": value, "unit": "correspondences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
            "config": workload_config(n1, n2, synth.SEED_TARGET),
            "method": {"timing": "per-step CUDA events on the library stream, summed; max over ranks",
                       "pairs_per_step": world},
            "icp_iters_per_s": args.steps * INNER_ITERS * world / dev_s,
            "build_ms": float(np.mean(build_ms)), "icp_ms": icp_avg_ms,
            "phases_ms": {"grid_build": float(np.mean(build_ms)), "spatial_order_of_source": float(np.mean(sort_ms)),
                          "iteration0_search_prepass": float(np.mean(pre_ms)), "persistent_kernel": kern_avg_ms,
                          "iteration1_search_kernel": float(np.mean(res_ms)),
                          "kernel_search_iterations": float(np.mean(it_search)),
                          "kernel_search_iteration_count": float(np.mean(n_search)),
                          "kernel_cached_iteration_us_median": cached_us},
            "natural": natural,
            "wall_s_timed_region": wall,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "correspondences/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 64, "ms_per_step": 1e3 * e2e_s / len(e2e_t), "samples": len(e2e_t),
                    "ms_min": 1e3 * min(e2e_t), "ms_max": 1e3 * max(e2e_t), "host_link_h2d_gbs": h2d_gbs},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "icp_persistent_kernel", "achieved": achieved,
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_CORR * INNER_ITERS * n2,
                         "launch_ms": kern_avg_ms, "launches_per_step": 2,
                         "achieved_incl_iteration1_search_kernel": ALG_BYTES_PER_CORR * INNER_ITERS * n2 / ((kern_avg_ms + float(np.mean(res_ms))) * 1e-3) / 1e9,
                         "share_of_step": kern_avg_ms * args.steps / (1e3 * dev_s) if world == 1 else None,
                         # a cached iteration alone (no search): algorithmic and streamed rate; at 1M the 64 MB of
                         # per-iteration streams are L2-resident, so this is L2 traffic, see "traffic" for the DRAM share
                         "cached_iteration": {"us": cached_us,
                                              "algorithmic_gbs": ALG_BYTES_PER_CORR * n2 / (cached_us * 1e-6) / 1e9,
                                              "algorithmic_frac_of_hbm_peak": ALG_BYTES_PER_CORR * n2 / (cached_us * 1e-6) / 1e9 / peak,
                                              "streamed_bytes_per_correspondence": STREAMED_BYTES_PER_CORR},
                         "note": "achieved = 48 B/correspondence x 50 iterations x n_source / duration of the "
                                 "icp_persistent_kernel (CUDA events around its two launches on the library stream, "
                                 "summed: iteration 0 | iterations 1..49); between them icp_research_kernel does the "
                                 "search of iteration 1 (every query searches there) at full occupancy -- "
                                 "achieved_incl_iteration1_search_kernel counts its time as well; the rest of a step is "
                                 "the grid build, the spatial ordering of the source and the iteration-0 search pre-pass "
                                 "(icp_seed_kernel); peak = measured copy bandwidth "
                                 "(MEASURED_PEAKS.json); traffic = dram read+write bytes of one launch (ncu)"},
        }
        # pose check against the oracle (full size is covered by tests -m gpu) + the CPU baseline, one thread like the reference
        if not args.no_cpu_baseline:
            from oracle import oracle_py as O
            t0 = time.perf_counter()
            sample_iters = 20                # ~10 s of single-thread CPU work at 1M
            o = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=sample_iters, force_iters=1))
            cpu_s = time.perf_counter() - t0
            g = ctx.icp_p2plane(h_t, h_n, h_s, P.icp_params(max_iter=sample_iters, force_iters=1))
            a, b = P.matrix2angle(g["T"]), P.matrix2angle(o["T"])
            line["pose_err_vs_oracle"] = {"rot_rad": float(np.abs(a - b).max()),
                                          "transl_m": float(np.abs(g["T"][:3, 3] - o["T"][:3, 3]).max()), "iters": sample_iters}
            line["cpu_baseline"] = {"value": sample_iters * n2 / cpu_s, "unit": "correspondences/s",
                                    "cores": 1, "kind": "port",
                                    "sample": f"{sample_iters} of {INNER_ITERS} inner iterations on the full "
                                              f"{n1}x{n2} pair incl. KD-tree build ({cpu_s:.1f} s); single "
                                              "thread, like the reference"}
        # secondary figures: the whole Piecewise_ICP outer loop (classification, inner ICP, bbox, DT schedule with stage-1
        # P75, transforms, VCM) at the centroid-level boundary -- 300k patches + 2.4M patch points, and BASELINE configs[0]
        # (the reference's shipped pair Epoch_001 -> Epoch_002, committed fixture) -- next to the oracle on the host cores
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                O.set_threads(threads)
                dp = synth.make_pair(300_000)
                pp = P.PairParams(dp["Res1"], dp["Res2"], dp["SVRes1"], dp["SVRes2"], dp["DTmin"])
                ctx.upload_pair(dp); ctx.piecewise_icp(pp, 1, 0.05)                     # warm-up
                ctx.upload_pair(dp)
                l0 = ctx.launch_count()
                g = ctx.piecewise_icp(pp, 1, 0.05)
                t0 = time.perf_counter()
                o = O.piecewise_icp(O.PairData(dp), 1, 0.05, O.icp_params(threads=threads))
                cpu_s = time.perf_counter() - t0
                a, b = P.matrix2angle(g["T"]), O.matrix2angle(o["T"])
                line["outer_loop"] = {"workload": "Piecewise_ICP outer loop, %d patches, %d patch points, DT 0.05 -> 0.004"
                                                  % (len(dp["ct2"]), len(dp["patch_pts2"])),
                                      "outer_iterations": int(g["n_outer"]), "device_ms": float(g["device_ms"]),
                                      "launches_per_outer_iteration": (ctx.launch_count() - l0) / max(1, int(g["n_outer"])),
                                      "inner_iterations": [int(st.icp_iters) for st in g["stats"]],
                                      "oracle_cpu_ms": 1e3 * cpu_s, "oracle_threads": threads,
                                      "DTseries_identical_to_oracle": bool(np.array_equal(g["DTseries"], o["DTseries"])),
                                      "pose_err_vs_oracle": {"rot_rad": float(np.abs(a - b).max()),
                                                             "transl_m": float(np.abs(g["T"][:3, 3] - o["T"][:3, 3]).max())}}
            except Exception as e:                                 # never lose the headline line to the extra
                line["outer_loop"] = {"error": str(e)[:200]}
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from conftest import load_refpair
                f = load_refpair(O.patch_stats)
                dp = f["pair"]
                pp = P.PairParams(dp["Res1"], dp["Res2"], dp["SVRes1"], dp["SVRes2"], dp["DTmin"])
                ctx.upload_pair(dp); ctx.piecewise_icp(pp, 1, f["DTinit"])
                ctx.upload_pair(dp)
                t0 = time.perf_counter()
                g = ctx.piecewise_icp(pp, 1, f["DTinit"])
                wall_ms = 1e3 * (time.perf_counter() - t0)
                t0 = time.perf_counter()
                o = O.piecewise_icp(O.PairData(dp), 1, f["DTinit"])
                cpu_s = time.perf_counter() - t0
                Tg = P.mat4_mul(P.mat4_mul(f["Sinv"], g["T"]), f["S"])           # back to the scans' frame, src/Registration.cpp:461
                a, b = P.matrix2angle(Tg), P.matrix2angle(f["T_recorded"].astype(np.float32))
                line["config0"] = {"workload": "BASELINE configs[0]: the reference's pair Epoch_001 -> Epoch_002 at the "
                                               "centroid-level boundary (%d / %d patches)" % (len(dp["ct1"]), len(dp["ct2"])),
                                   "outer_iterations": int(g["n_outer"]), "device_ms": float(g["device_ms"]), "wall_ms": wall_ms,
                                   "oracle_cpu_ms_one_thread": 1e3 * cpu_s,
                                   "pose_err_vs_recorded_result": {"rot_rad": float(np.abs(a - b).max()),
                                                                   "transl_m": float(np.abs(Tg[:3, 3] - f["T_recorded"][:3, 3]).max())}}
            except Exception as e:
                line["config0"] = {"error": str(e)[:200]}
            # secondary figure: PCpreprocessing (pcl::VoxelGrid + StatisticalOutlierRemoval, k = 14) on the 1M-point cloud,
            # device (events around voxel grid + grid build + 14-NN kernel) next to the CPU restatement (one thread)
            try:
                c = d["ct1"]
                for _ in range(2):
                    out = ctx.preprocess(c, 0.05, 14, 5.0)
                dev_ms, knn_ms = ctx.last_device_ms(), ctx.last_knn_kernel_ms()
                t0 = time.perf_counter()
                ref = O.preprocess(c, 0.05, 14, 5.0)
                cpu_s = time.perf_counter() - t0
                line["preprocess"] = {"workload": "PCpreprocessing(leaf 0.05, k 14, 5 sigma), %d -> %d points" % (len(c), len(out)),
                                      "device_ms": float(dev_ms), "knn_kernel_ms": float(knn_ms), "cpu_ms": 1e3 * cpu_s,
                                      "identical_to_cpu": bool(out.shape == ref.shape and (out == ref).all())}
            except Exception as e:
                line["preprocess"] = {"error": str(e)[:200]}
    ctx.close()

    # ---- BASELINE configs[3]: the epoch-sharded 4D series (every N) ------------------------------
    if not args.no_config4:
        if world > 1:                # an exception on one rank must not leave the others in a collective: let it propagate
            c4 = config4_leg(P, synth, torch, dist, local_rank, rank, world)
        else:
            try:
                c4 = config4_leg(P, synth, torch, dist, local_rank, rank, world)
            except Exception as e:
                c4 = {"error": str(e)[:300]}
        if line is not None:
            line["config4"] = c4
    if line is not None:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
