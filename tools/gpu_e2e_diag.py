"""where the end-to-end step of bench.py spends its time beyond the solve kernel (run on the GPU box)"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.packing import PackedModel

fixture, kind, B, H, iters, integ, lr, scaling, desc = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4_rk4"]
sd = bench.load_fixture(fixture)
c = bench.cost_for(kind)
spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
mpc = BatchedMPC(pk, H, 0.02, spec, integrator=integ, lr=lr, iters=iters, return_mode="last")
x0_host = bench.make_inputs(B, kind, 7).pin_memory()
x0 = x0_host.cuda()
U_host = torch.empty((B, H, 1), dtype=torch.float32).pin_memory()
c_host = torch.empty((B,), dtype=torch.float32).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    mpc.solve(x0)
torch.cuda.synchronize()

def ev():
    return torch.cuda.Event(enable_timing=True)

# (a) back to back, events
evs = []
for _ in range(3):
    flush.zero_(); a, b = ev(), ev(); a.record(); mpc.solve(x0); b.record(); evs.append((a, b))
torch.cuda.synchronize()
print("back-to-back, events around solve:      ", ["%.1f" % a.elapsed_time(b) for a, b in evs], flush=True)
# (b) each step after a synchronize (idle gap), events + wall clock
for gap in (0.0, 0.05, 0.5):
    res = []
    for _ in range(3):
        flush.zero_(); torch.cuda.synchronize(); time.sleep(gap)
        t0 = time.perf_counter(); a, b = ev(), ev(); a.record(); o = mpc.solve(x0); b.record()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        res.append("ev %.1f wall %.1f (host issue %.1f)" % (a.elapsed_time(b), 1e3 * (t2 - t0), 1e3 * (t1 - t0)))
    print("after sync + %.2f s idle:" % gap, res, flush=True)
# (c) the e2e step of bench.py, segment by segment
for _ in range(3):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    xd = x0_host.to("cuda", non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    o = mpc.solve(xd); torch.cuda.synchronize(); t2 = time.perf_counter()
    U_host.copy_(o["U"], non_blocking=True); c_host.copy_(o["best_cost"], non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("e2e segments: H2D %.2f ms, solve %.1f ms, D2H %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
# (d) as bench.py does it (one synchronize at the end)
for _ in range(3):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    xd = x0_host.to("cuda", non_blocking=True)
    o = mpc.solve(xd)
    U_host.copy_(o["U"], non_blocking=True); c_host.copy_(o["best_cost"], non_blocking=True); torch.cuda.synchronize()
    print("e2e as bench.py: %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
