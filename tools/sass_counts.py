"""SASS opcode histogram per kernel of libphnn_mpc.so (cuobjdump -sass): the evidence that the tcgen05 / bulk-copy
instructions are in the shipped binary.  python tools/sass_counts.py [lib] > profiles/rNN_sass_counts.txt"""
import collections, os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(R, "phnn_mpc_b200", "libphnn_mpc.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDS", "STS",
       "LDG", "STG", "HMMA", "SHFL", "BAR", "NANOSLEEP", "USETMAXREG"]
kern = None; hist = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(phnn::KParams\)|void |phnn::|\(int\)", "", kern)
        hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern: hist[kern][m.group(1)] += 1
print("SASS opcode counts per kernel of %s (static instruction counts; cuobjdump -sass, sm_100a)" % os.path.relpath(lib, R))
print("%-44s %7s " % ("kernel", "total") + " ".join("%9s" % k for k in KEY))
for k, h in sorted(hist.items()):
    print("%-44s %7d " % (k[:44], sum(h.values())) + " ".join("%9d" % h.get(x, 0) for x in KEY))
print("\nUTCHMMA = tcgen05.mma kind::f16/tf32, LDTM/STTM = tcgen05.ld/st (tensor memory), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk "
      "(1-D TMA bulk copy), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg")
