"""Print selected metrics of an exported raw page (ncu -i rep --page raw --csv > file.csv): python tools/ncu_raw_q.py file.csv regex..."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
pats = [re.compile(p) for p in sys.argv[2:]]
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if any(p.search(h) for p in pats):
        print("%-90s %-12s %s" % (h, u, v))
