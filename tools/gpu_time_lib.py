"""times bench workloads with the library named by PHNN_MPC_LIB (timing-only experiment builds): python tools/gpu_time_lib.py [workload ...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.packing import PackedModel
sd = bench.load_fixture("cartpole_h256")
c = bench.cost_for("phnn")
spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
for wl in (sys.argv[1:] or ["small", "cfg4_rk4"]):
    fixture, kind, B, H, iters, integ, lr, scaling, desc = bench.WORKLOADS[wl]
    x0 = bench.make_inputs(B, kind, 7).cuda()
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
    mpc = BatchedMPC(pk, H, 0.02, spec, integrator=integ, lr=lr, iters=iters, return_mode="last")
    for _ in range(2):
        mpc.solve(x0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.solve(x0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    out = mpc.solve(x0)
    torch.cuda.synchronize()
    U = out["U"].double()
    print("%s %-9s %.2f ms (%.1f k solves/s)  checksum U %.9e |U| %.9e" % (os.path.basename(os.environ.get("PHNN_MPC_LIB", "default")), wl, np.median(ts),
          B / np.median(ts), U.sum().item(), U.abs().sum().item()), flush=True)
