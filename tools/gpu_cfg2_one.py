"""BASELINE cfg2 rollout (4096 x H=100, RK4, pendulum pHNN) on its default route, three launches (profiling target)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench as B
from phnn_mpc_b200.batched import rollout
from phnn_mpc_b200.packing import PackedModel
dev = torch.device("cuda", 0)
sdp = B.load_fixture("pendulum")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sdp.items()}, "phnn", device=dev)
if len(sys.argv) > 1 and sys.argv[1] == "latency":
    pk.set_option("tensor_fwd_min_batch", 0)
g = torch.Generator().manual_seed(1)
Bp, Tp = 4096, 100
xp = torch.stack([(torch.rand(Bp, generator=g) * 2 - 1) * np.pi, torch.rand(Bp, generator=g) * 2 - 1], 1).to(dev)
Up = (torch.rand(Bp, Tp, 1, generator=g) * 4 - 2).to(dev)
for _ in range(3):
    rollout(pk, xp, Up, 0.05, "rk4")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); rollout(pk, xp, Up, 0.05, "rk4"); e1.record(); torch.cuda.synchronize()
print("cfg2 rollout %.3f ms" % e0.elapsed_time(e1))
