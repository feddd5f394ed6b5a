"""Per-phase cycle breakdown of the tcgen05 kernel (needs a -DPHNN_TC_PROFILE build)."""
import ctypes, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import load_golden
from phnn_mpc_b200 import ops, _lib
from phnn_mpc_b200.packing import PackedModel
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
z, sd = load_golden("cartpole_h256")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
pk.set_option("tensor_min_batch", 0)
if os.environ.get("PHNN_TMODE"): pk.set_option("tensor_mode", int(os.environ["PHNN_TMODE"]))
dbg = torch.zeros(48, dtype=torch.int64, device="cuda")
L = _lib.lib(); L.phnn_debug_set_buffer.argtypes = [ctypes.c_void_p]; L.phnn_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
H, iters = 50, 2
g = torch.Generator().manual_seed(7)
x0 = ((torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).cuda()
mpc = BatchedMPC(pk, H, 0.02, CostSpec.make(4, 1, [10.0, 200.0, 1.0, 10.0], [0.01], None, -15.0, 15.0), integrator="rk4", lr=0.015, iters=iters)
for _ in range(2):
    mpc.solve(x0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mpc.solve(x0); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
npair = iters * H * 4
names = ["fwdA:a1+Rnet", "wait acc z2", "fwdB:a2,delta2", "wait acc g1", "fwdC:gradH", "exchanges", "adjA3:da1+xbar(g1)+R/2", "adj wait dz2",
         "adjB3:e2+R/2", "adj wait dg1", "adjC4", "-", "-", "-", "-", "outside evals"]
d = dbg.cpu().numpy()[:32].reshape(2, 16)
print("B=%d  %.2f ms  -> %.0f cycles per (fwd+adj) pair @1.965GHz" % (B, ms, ms * 1e-3 * 1.965e9 / npair))
for i, n in enumerate(names):
    print("  %-18s thread0 %8.0f   thread255 %8.0f   cycles/pair" % (n, d[0, i] / npair, d[1, i] / npair))
print("  total              thread0 %8.0f" % (d[0].sum() / npair))
mw = dbg.cpu().numpy()[32:34]
print("  MMA issuer waits per pair: operand A (element threads) %.0f, operand B (weight stream) %.0f" % (mw[0] / npair, mw[1] / npair))
