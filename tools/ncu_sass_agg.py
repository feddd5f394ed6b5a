"""Aggregate the SASS source page of an ncu report (ncu -i rep --page source --csv --print-source sass):
per opcode executed warp-instructions, stall samples by reason, shared wavefronts.  python tools/ncu_sass_agg.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
op_inst = collections.Counter(); op_samp = collections.Counter(); op_wave = collections.Counter()
reason_tot = collections.Counter(); op_reason = collections.defaultdict(collections.Counter)
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    w = int(r[ix["L1 Wavefronts Shared"]] or 0)
    op_inst[op] += n; op_samp[op] += s; op_wave[op] += w; tot_i += n; tot_s += s
    for re_ in reasons:
        v = int(r[ix[re_]] or 0)
        reason_tot[re_] += v; op_reason[op][re_] += v
print("total warp-inst %d  samples %d" % (tot_i, tot_s))
print("stall reasons (all samples):", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot_s) for k, v in reason_tot.most_common(12)))
print("%-28s %8s %8s %10s  top reasons" % ("opcode", "inst%", "samp%", "smem wave%"))
tw = sum(op_wave.values()) or 1
for op, n in op_inst.most_common(45):
    top = ", ".join("%s %.1f" % (k[6:], 100.0 * v / tot_s) for k, v in op_reason[op].most_common(3) if v)
    print("%-28s %8.2f %8.2f %10.2f  %s" % (op, 100.0 * n / tot_i, 100.0 * op_samp[op] / tot_s, 100.0 * op_wave[op] / tw, top))
