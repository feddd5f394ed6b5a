"""Aggregate the correlated CUDA+SASS source page (ncu -i rep --page source --csv --print-source cuda,sass) per source
line: executed warp-instructions, stall samples, shared-memory wavefronts.  python tools/ncu_line_agg.py file.csv [topN]"""
import csv, sys, collections
def num(x):
    try: return int(x)
    except ValueError: return 0
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None; ix = None
inst = collections.Counter(); samp = collections.Counter(); wave = collections.Counter(); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; iA = hdr.index("Address"); continue
    if hdr is None or len(r) < len(hdr): continue
    if not r[0].strip().isdigit(): continue
    key = (cur, int(r[0]))
    text[key] = r[1].strip()[:110]
    inst[key] += num(r[ix["Instructions Executed"]])
    samp[key] += num(r[ix["# Samples"]])
    wave[key] += num(r[ix["L1 Wavefronts Shared"]])
ti, ts, tw = sum(inst.values()), sum(samp.values()), sum(wave.values()) or 1
print("total inst %d samples %d" % (ti, ts))
for key, s in samp.most_common(top):
    print("%-22s %5d  inst %5.2f%% samp %5.2f%% wave %5.2f%%  %s" % (key[0], key[1], 100.0 * inst[key] / ti, 100.0 * s / ts, 100.0 * wave[key] / tw, text[key]))
