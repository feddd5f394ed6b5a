"""single-instance solve through the latency kernel (profiling target)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch
import bench as B
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.packing import PackedModel
dev = torch.device("cuda", 0)
sdm = B.load_fixture("cartpole_h128")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sdm.items()}, "phnn", device=dev)
cost = B.cost_for("phnn")
spec = CostSpec.make(4, 1, cost["Q"], cost["R"], None, cost["u_min"], cost["u_max"])
x = B.make_inputs(1, "phnn", 3).to(dev)
mpc = BatchedMPC(pk, 20, 0.02, spec, integrator="euler", lr=0.015, iters=30)
for _ in range(3):
    mpc.solve(x)
torch.cuda.synchronize()
