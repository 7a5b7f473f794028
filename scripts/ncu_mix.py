"""Instruction mix, lane efficiency and the main counters of one kernel in an ncu report (read on the CPU box).
python scripts/ncu_mix.py <report.ncu-rep> [top_n_lines]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "launch__registers_per_thread",
        "smsp__average_warp_latency_issue_stalled", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu",
        "smsp__average_warps_issue_stalled", "dram__bytes_read.sum", "launch__occupancy_limit", "achieved_occupancy",
        "sm__throughput.avg.pct", "smsp__warp_issue_stalled"]
for h, u, v in zip(hdr, units, r):
    if any(w in h for w in want) and v not in ("", "0", "n/a"):
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if x and x[0] == "Address")
hdr = rows[hi]; data = rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
def num(x):
    try: return int(x)
    except ValueError: return 0
ti = sum(num(x[ix["Instructions Executed"]]) for x in data)
tt = sum(num(x[ix["Thread Instructions Executed"]]) for x in data)
ts = sum(num(x[ix["# Samples"]]) for x in data)
print(f"warp instructions {ti}, thread instructions {tt}, lanes/inst {tt / max(ti, 1):.2f}, samples {ts}")
c = Counter(); s = Counter()
for x in data:
    o = [t for t in x[ix["Source"]].split() if not t.startswith("@")]
    if not o: continue
    m = o[0].split(".")[0]
    c[m] += num(x[ix["Instructions Executed"]]); s[m] += num(x[ix["# Samples"]])
for m, n in c.most_common(18):
    print(f"  {m:10s} {n:11d} {100 * n / ti:5.1f}%   samples {100 * s[m] / max(ts, 1):5.1f}%")
if topn:
    print("hottest lines by samples:")
    for x in sorted(data, key=lambda x: -num(x[ix["# Samples"]]))[:topn]:
        print(f"  {num(x[ix['# Samples']]):6d} {num(x[ix['Instructions Executed']]):9d} {x[ix['Source']][:90]}")
