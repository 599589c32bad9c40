"""SASS-level view of an .ncu-rep around the tensor-core instructions: stall samples of every instruction with
at least `min_samples` samples plus all UTC* (tcgen05) instructions, in program order."""
import csv, subprocess, sys
path = sys.argv[1]
min_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
i_s, i_i = H.index("# Samples"), H.index("Instructions Executed")
ins = [(r[1].strip(), int(r[i_s]), int(r[i_i])) for r in rows[hdr + 1:] if len(r) > i_i and r[i_s].isdigit()]
tot = sum(x[1] for x in ins)
print(f"{path}: {len(ins)} SASS instructions, {tot} samples")
for i, (txt, s, n) in enumerate(ins):
    if s >= min_samples or "UTCHMMA" in txt or "UTCBAR" in txt:
        print(f"{i:6d} {s:6d} {100*s/tot:5.1f}% {n:10d}  {txt[:100]}")
