"""Per-source-line warp-stall samples of an .ncu-rep (needs -lineinfo + --import-source on): prints the
hottest CUDA source lines with their share of all samples and of the instructions executed.
    python tools/ncu_hot_lines.py report.ncu-rep [top [launch index]]"""
import csv, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sel = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []   # one launch of a multi-launch report
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", *sel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
i_samp, i_inst = H.index("# Samples"), H.index("Instructions Executed")
fname = ""
lines = []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > i_inst and r[0].isdigit():
        try:
            lines.append((int(r[i_samp]), int(r[i_inst]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_s = sum(l[0] for l in lines) or 1
tot_i = sum(l[1] for l in lines) or 1
print(f"{path}: {tot_s} samples, {tot_i} warp instructions")
for s, i, f, n, src in sorted(lines, reverse=True)[:top]:
    print(f"  {100*s/tot_s:5.1f}% samp {100*i/tot_i:5.1f}% inst  {f}:{n}  {src[:110]}")
