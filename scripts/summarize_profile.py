"""Turns the captures of scripts/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries
under profiles/: launch shares, key ncu metrics of the dominant kernel, DRAM traffic per launch."""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles"); src = os.path.join(ROOT, "gpurun_out")

# (1) launch shares
rows = [r for r in csv.reader(l for l in open(os.path.join(src, f"{tag}_launches.csv")) if not l.startswith("==")) if r]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    tot[r[ik]] += ms; cnt[r[ik]] += 1
allms = sum(tot.values())
with open(os.path.join(out, f"{tag}_launch_shares.txt"), "w") as f:
    f.write("# launch list shares: ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 "
            "--warmup 3 --no-cpu-baseline (per-launch times are cold-cache and serialised; shares, not absolutes)\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        f.write(f"{tot[k]:10.3f} ms {cnt[k]:4d}x {100 * tot[k] / allms:5.1f}%  {k[:70]}\n")
subprocess.run(["cp", os.path.join(src, f"{tag}_launches.csv"), os.path.join(out, f"{tag}_launches.csv")])

# (2) key metrics of the full capture
raw = subprocess.run(["ncu", "-i", os.path.join(src, f"{tag}_icp_bench.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ikn, idur = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
captured = [r for r in rows[2:] if len(r) == len(hdr)]
pers = [r for r in captured if "icp_persistent" in r[ikn]]
vals = max(pers, key=lambda r: float(r[idur].replace(",", "")))      # the long launch (iterations 1..49)
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_stores", "smsp__inst_executed_op_ldgsts.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_global_loads", "sass__inst_executed_shared_loads"]
d = {}
with open(os.path.join(out, f"{tag}_ncu_icp.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 (1M source x 1M "
            "target, 50 forced inner iterations): the kernels of the inner loop of one step; the full list of metrics is for "
            "the long launch of icp_persistent_kernel (iterations 1..49)\n")
    for r in captured:
        f.write(f"#   {r[ikn][:60]:60s} {r[idur]} {units[idur]}, dram read {r[hdr.index('dram__bytes_read.sum')]} "
                f"{units[hdr.index('dram__bytes_read.sum')]}, write {r[hdr.index('dram__bytes_write.sum')]} {units[hdr.index('dram__bytes_write.sum')]}, "
                f"lanes/inst {r[hdr.index('smsp__thread_inst_executed_per_inst_executed.ratio')]}, issue active "
                f"{r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]} %\n")
    for k in want:
        if k in hdr:
            i = hdr.index(k); f.write(f"{k} [{units[i]}] = {vals[i]}\n"); d[k] = (vals[i], units[i])
def to_bytes(v, u):
    v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
ird, iwr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")      # both launches of the persistent kernel
tr = sum(to_bytes(r[ird], units[ird]) + to_bytes(r[iwr], units[iwr]) for r in pers)
json.dump({"kernel": "icp_persistent_kernel", "config": "1M x 1M, 50 forced inner iterations (bench.py)",
           "dram_bytes_per_launch": tr, "source": f"profiles/{tag}_ncu_icp.txt"},
          open(os.path.join(out, "traffic.json"), "w"))
print("traffic per launch", tr / 1e6, "MB")
bl = os.path.join(src, f"{tag}_bench_line.json")
if os.path.exists(bl) and os.path.getsize(bl):
    subprocess.run(["cp", bl, os.path.join(out, f"{tag}_bench_line.json")])

# (3) the 10M capture (BASELINE configs[4]), when present
p10 = os.path.join(src, f"{tag}_icp_10m.ncu-rep")
if os.path.exists(p10):
    raw = subprocess.run(["ncu", "-i", p10, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(out, f"{tag}_ncu_icp_10m.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none, icp_persistent_kernel, the long launch (iterations 1..49) of the 3rd run of python scripts/prof_icp_only.py 10000000 50 "
                "(10M source x 10M target, 50 forced inner iterations; BASELINE configs[4])\n")
        for k in want:
            if k in hdr:
                i = hdr.index(k); f.write(f"{k} [{units[i]}] = {vals[i]}\n")
