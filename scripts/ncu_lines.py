"""Per-source-line totals (warp instructions, lanes per instruction, stall samples) of a kernel in an ncu report taken
with --import-source on from a -lineinfo build.  python scripts/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; lines = []
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":   # a source line with the totals of its SASS
        lines.append((fname, r[0], r[1], num(r[ix["Instructions Executed"]]), num(r[ix["Thread Instructions Executed"]]), num(r[ix["# Samples"]])))
ti = sum(l[3] for l in lines); ts = sum(l[5] for l in lines)
print(f"warp instructions {ti}, samples {ts}")
for f, ln, src, wi, th, sm in sorted(lines, key=lambda l: -l[3])[:topn]:
    print(f"{f}:{ln:>4s} {wi:10d} {100 * wi / max(ti, 1):5.1f}%  lanes {th / max(wi, 1):4.1f}  samples {100 * sm / max(ts, 1):5.1f}% | {src.strip()[:100]}")
