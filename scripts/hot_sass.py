"""Per-instruction stall samples of an ncu capture taken with --import-source on: the hottest SASS instructions and the
share of the samples by code region.  python scripts/hot_sass.py <file.ncu-rep> <out.txt> '<title>' [ncu import filters, e.g. --launch-skip 3 --launch-count 1]"""
import csv, subprocess, sys
from collections import defaultdict
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
pick = sys.argv[4] if len(sys.argv) > 4 else ""      # substring of the kernel name; of several matches the one with the most samples
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
best = None
for a, b in zip(starts[:-1], starts[1:]):
    if pick not in rows[a][1]: continue
    h = rows[a + 1]; dd = [r for r in rows[a + 2:b] if len(r) == len(h)]
    n = sum(int(r[h.index("# Samples")]) for r in dd)
    if best is None or n > best[0]: best = (n, h, dd)
hdr, data = best[1], best[2]
iS, iE, isrc, iA = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Address")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
def top(r):
    s = sorted(((int(r[i]), hdr[i][6:]) for i in stall if r[i] not in ("", "0")), reverse=True)
    return " ".join(f"{h}:{v}" for v, h in s[:3])
with open(out, "w") as f:
    f.write(f"# {title}\n# SASS instructions with the most warp-stall samples (share of all {tot} samples), executed-instruction counts, top stall reasons\n")
    for r in sorted(data, key=lambda r: -int(r[iS]))[:25]:
        f.write(f"{100 * int(r[iS]) / tot:5.1f}%  {int(r[iE]):10d} exec  {r[iA][-6:]}  {r[isrc].strip()[:64]:64s} {top(r)}\n")
    agg = defaultdict(int)
    for r in data:
        for i in stall:
            if r[i] not in ("", "0"): agg[hdr[i][6:]] += int(r[i])
    f.write("# all samples by stall reason: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]) + "\n")
print(open(out).read())
