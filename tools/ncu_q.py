"""Print selected metrics of an .ncu-rep (first profiled launch): python tools/ncu_q.py rep regex..."""
import csv, re, subprocess, sys
rep, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if any(p.search(h) for p in pats):
        print("%-90s %-12s %s" % (h, u, v))
