"""Single-instance solve latency breakdown (debug aid)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import pack_of
from phnn_mpc_b200.dropin.pHNN import pHNN
from phnn_mpc_b200.dropin.mpc_controller import MPCController
torch.manual_seed(0)
m = pHNN(os.path.join(R, "configs", "cartpole_phnn.yaml"))
c = MPCController(m, 20, 0.02, [10.0, 200.0, 1.0, 10.0], 0.01, [0, 0, 0, 0], -15.0, 15.0, lr=0.015, max_iterations=30)
s = np.array([0.0, 0.1, 0.0, 0.0])
for _ in range(3): c.compute_control(s)
t0 = time.perf_counter()
for _ in range(10): c.compute_control(s)
print("compute_control: %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
t0 = time.perf_counter()
for _ in range(100): pk = pack_of(m)
print("pack_of (cached): %.3f ms" % ((time.perf_counter() - t0) / 100 * 1e3))
eng = c._engine()
x0 = torch.tensor(s, dtype=torch.float32).reshape(1, 4).cuda()
U0 = torch.zeros(1, 20, 1, device="cuda")
args = eng.cost.op_args()
for B in (1, 32, 148 * 32, 4096):
    xb, Ub = x0.repeat(B, 1), U0.repeat(B, 1, 1)
    for _ in range(2): ops.mpc_solve(pk.handle, xb, Ub, 0.02, 0, *args, 0.015, 0.9, 0.999, 1e-8, 30, 0, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mpc_solve(pk.handle, xb, Ub, 0.02, 0, *args, 0.015, 0.9, 0.999, 1e-8, 30, 0, False); e1.record(); torch.cuda.synchronize()
    print("B=%d kernel+op: %.2f ms (device events)" % (B, e0.elapsed_time(e1)))
