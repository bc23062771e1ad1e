"""Per-SASS-line stall breakdown of an ncu report: python tools/ncu_lines.py rep.ncu-rep idx [idx ...] (line indices as printed by ncu_phase.py)"""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]
ia, isamp = h2.index("Source"), h2.index("# Samples")
cols = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
for a in sys.argv[2:]:
    lo, _, hi = a.partition("-")
    for i in range(int(lo), int(hi or lo) + 1):
        r = data[i]
        d = {c[6:]: int(r[h2.index(c)]) for c in cols if int(r[h2.index(c)]) > 0}
        print(i, r[isamp], r[ia].strip()[:70], dict(sorted(d.items(), key=lambda kv: -kv[1])[:4]))
