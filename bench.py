#!/usr/bin/env python
"""bench.py -- headline benchmark of the Piecewise-ICP inner registration loop on B200.

Workload (BASELINE.json configs[1]): pairwise, 1M-centroid synthetic planar-patch pair, 50 inner
point-to-plane ICP iterations (convergence test evaluated, not allowed to stop the loop).
One "step" = device build of the target grid + the 50-iteration inner loop over all source
centroids (per-pair grid build included, one-time uploads excluded).

  value : correspondences/s with inputs already resident in HBM (CUDA events on the library stream)
  e2e   : the same metric through the host-buffer C-ABI call pwicp_icp_p2plane (the call shape of
          P2PICPwithPatchNormal), H2D of both clouds and D2H of the 4x4 inside the timed region
  N>1   : independent pairs ("epochs") sharded one per rank, no data-path collective; the per-pair
          384-byte result records are all-gathered once at the end; weak scaling.

`--impl reference` times the CPU oracle (oracle/, the restatement of the reference's PCL path; the
reference itself cannot be built here) on the host cores, single thread like the
reference, on a bounded sample of the same workload.
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
METRIC = "correspondences/s/GPU (ICP iters/s on 1M-pt pair; pose err vs ref)"


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


def pose_error(T, T_ref, matrix2angle):
    a, b = matrix2angle(T), matrix2angle(T_ref)
    return float(np.abs(a - b).max()), float(np.abs(T[:3, 3] - T_ref[:3, 3]).max())


def run_reference(args, rank, world):
    """CPU arm: the oracle's inner loop on the host cores.  The reference itself is single-threaded
    (no OpenMP / threads anywhere in its tree); the figure reported here lets it use every host thread
    for the independent NN queries and row terms (sums stay sequential), and states the single-thread
    figure next to it."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    from pwicp_b200 import synth
    d = synth.make_pair(N_CENTROIDS, with_clouds=False)
    n1, n2 = len(d["ct1"]), len(d["ct2"])
    threads = os.cpu_count() or 1
    sample_iters = 4
    prm = O.icp_params(max_iter=sample_iters, force_iters=1, threads=threads)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.icp(d["ct1"], d["nrm1"], d["ct2"], prm)        # tree build + sample_iters iterations
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = len(times) * sample_iters * n2 / tot
    t0 = time.perf_counter()
    O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=sample_iters, force_iters=1))
    single = sample_iters * n2 / (time.perf_counter() - t0)
    sample = (f"{sample_iters} of {INNER_ITERS} inner iterations on the full {n1}x{n2} pair, KD-tree "
              f"build (serial) included, per step; NN queries on {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "correspondences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
        "config": {"workload": "pairwise 1M-centroid synthetic planar-patch pair, 50 inner ICP iterations",
                   "n_target": n1, "n_source": n2, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "correspondences/s", "cores": threads, "kind": "port",
                         "sample": sample, "single_thread_value": single,
                         "note": "the reference is single-threaded; kind=port because PCL/Eigen/Boost are absent "
                                 "(DESIGN.md section 7)"},
        "e2e": {"value": value, "unit": "correspondences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_CENTROIDS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    build_ms, icp_ms, kern_ms, corr = [], [], [], 0
    last = None
    for _ in range(args.steps):
        b_ms, r = step_resident()
        build_ms.append(b_ms); icp_ms.append(r["device_ms"]); kern_ms.append(r["kernel_ms"]); corr += r["correspondences"]
        last = r
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count() - launches0 - args.steps    # minus the L2-flush fills
    clocks = sampler.stop(epoch0, time.time())
    dev_s = (sum(build_ms) + sum(icp_ms)) / 1e3

    # ---- e2e arm: host buffers through the reference-shaped call ---------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    h_t, h_n, h_s = pin(d["ct1"]), pin(d["nrm1"]), pin(d["ct2"])
    ctx.icp_p2plane(h_t, h_n, h_s, prm)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_t = []
    for _ in range(max(3, args.steps // 2)):
        ctx.flush_l2()
        t0 = time.perf_counter()
        r2 = ctx.icp_p2plane(h_t, h_n, h_s, prm)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = sum(e2e_t)
    e2e_corr = len(e2e_t) * INNER_ITERS * n2
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

    if rank == 0:
        peak, peak_kind = measured_hbm_peak()
        traffic, traffic_src = profiled_traffic()
        icp_avg_ms = float(np.mean(icp_ms))
        kern_avg_ms = float(np.mean(kern_ms))
        achieved = ALG_BYTES_PER_CORR * INNER_ITERS * n2 / (kern_avg_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "correspondences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
            "config": {"workload": "pairwise 1M-centroid synthetic planar-patch pair, 50 inner ICP iterations",
                       "n_target": n1, "n_source": n2, "inner_iters": INNER_ITERS,
                       "step": "device grid build over the target + 50 forced inner iterations",
                       "l2": "flushed (384 MiB fill) between timed steps",
                       "timing": "per-step CUDA events on the library stream, summed; max over ranks",
                       "pairs_per_step": world, "seed": synth.SEED_TARGET},
            "icp_iters_per_s": args.steps * INNER_ITERS * world / dev_s,
            "build_ms": float(np.mean(build_ms)), "icp_ms": icp_avg_ms,
            "wall_s_timed_region": wall,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "correspondences/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 64, "ms_per_step": 1e3 * e2e_s / len(e2e_t)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "icp_persistent_kernel", "achieved": achieved,
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_CORR * INNER_ITERS * n2,
                         "launch_ms": kern_avg_ms, "share_of_step": kern_avg_ms * args.steps / (1e3 * dev_s) if world == 1 else None,
                         # what the kernel actually moves per correspondence (four 16-byte reads + one write; at 1M the
                         # streams are L2-resident, so this is L2 traffic, see "traffic" for the DRAM share)
                         "streamed": {"bytes_per_correspondence": 80,
                                      "gbs": achieved * 80.0 / ALG_BYTES_PER_CORR, "frac_of_hbm_peak": achieved * 80.0 / ALG_BYTES_PER_CORR / peak},
                         "note": "achieved = 48 B/correspondence x 50 iterations x n_source / average duration of "
                                 "the icp_persistent_kernel launch (CUDA events around the launch on the library "
                                 "stream); the rest of a step is the grid build, the Morton sort of the source and "
                                 "the iteration-0 search pre-pass (icp_seed_kernel); peak = measured copy bandwidth "
                                 "(MEASURED_PEAKS.json); traffic = dram read+write bytes of one launch (ncu)"},
        }
        # pose check against the oracle on a small pair (full size is covered by tests -m gpu)
        if not args.no_cpu_baseline:
            from oracle import oracle_py as O
            t0 = time.perf_counter()
            sample_iters = 20                # ~10 s of single-thread CPU work at 1M
            o = O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=sample_iters, force_iters=1))
            cpu_s = time.perf_counter() - t0
            g = ctx.icp_p2plane(h_t, h_n, h_s, P.icp_params(max_iter=sample_iters, force_iters=1))
            rot, tr = pose_error(g["T"], o["T"], P.matrix2angle)
            line["pose_err_vs_oracle"] = {"rot_rad": rot, "transl_m": tr, "iters": sample_iters}
            line["cpu_baseline"] = {"value": sample_iters * n2 / cpu_s, "unit": "correspondences/s",
                                    "cores": 1, "kind": "port",
                                    "sample": f"{sample_iters} of {INNER_ITERS} inner iterations on the full "
                                              f"{n1}x{n2} pair incl. KD-tree build ({cpu_s:.1f} s); single "
                                              "thread, like the reference"}
            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            O.icp(d["ct1"], d["nrm1"], d["ct2"], O.icp_params(max_iter=sample_iters, force_iters=1, threads=threads))
            mt_s = time.perf_counter() - t0
            line["cpu_baseline_mt"] = {"value": sample_iters * n2 / mt_s, "unit": "correspondences/s",
                                       "cores": threads, "kind": "port",
                                       "sample": f"same sample, NN queries and row terms on {threads} threads "
                                                 f"({mt_s:.1f} s); not what the reference does"}
        # secondary figure: the whole Piecewise_ICP outer loop (classification, inner ICP, bbox, DT schedule
        # with stage-1 P75, transforms, VCM) at the centroid-level boundary, 300k patches + 2.4M patch points
        if world == 1 and not args.no_cpu_baseline:
            try:
                dp = synth.make_pair(300_000)
                ctx.upload_pair(dp)
                pp = P.PairParams(dp["Res1"], dp["Res2"], dp["SVRes1"], dp["SVRes2"], dp["DTmin"])
                ctx.piecewise_icp(pp, 1, 0.05)                     # warm-up
                ctx.upload_pair(dp)
                g = ctx.piecewise_icp(pp, 1, 0.05)
                line["outer_loop"] = {"workload": "Piecewise_ICP outer loop, %d patches, %d patch points, DT 0.05 -> 0.004"
                                                  % (len(dp["ct2"]), len(dp["patch_pts2"])),
                                      "outer_iterations": int(g["n_outer"]), "device_ms": float(g["device_ms"]),
                                      "inner_iterations": [int(st.icp_iters) for st in g["stats"]]}
            except Exception as e:                                 # never lose the headline line to the extra
                line["outer_loop"] = {"error": str(e)[:200]}
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
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
