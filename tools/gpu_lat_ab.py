"""Timing of the latency kernel's jobs (A/B across library builds via PHNN_MPC_LIB): cfg2 rollout, small-batch solves."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench as B
from phnn_mpc_b200.batched import BatchedMPC, CostSpec, rollout
from phnn_mpc_b200.packing import PackedModel
dev = torch.device("cuda", 0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


sd = B.load_fixture("pendulum")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn", device=dev)
g = torch.Generator().manual_seed(1)
Bn, T = 4096, 100
x0 = torch.stack([(torch.rand(Bn, generator=g) * 2 - 1) * np.pi, torch.rand(Bn, generator=g) * 2 - 1], 1).to(dev)
U = (torch.rand(Bn, T, 1, generator=g) * 4 - 2).to(dev)
ms = timed(lambda: rollout(pk, x0, U, 0.05, "rk4"))
print("cfg2 pendulum rollout 4096x100 rk4: %.3f ms  %.1f M inst-steps/s" % (ms, Bn * T / ms / 1e3))
for name, kind, H, it, lr, nb in (("cartpole_h128", "phnn", 20, 30, 0.015, (1, 64, 512)), ("canonical", "canonical", 10, 50, 0.03, (1, 64, 512, 888))):
    sdm = B.load_fixture(name)
    pkm = PackedModel({k: torch.from_numpy(v) for k, v in sdm.items()}, kind, device=dev)
    cost = B.cost_for(kind)
    spec = CostSpec.make(4, 1, cost["Q"], cost["R"], None, cost["u_min"], cost["u_max"])
    for nbi in nb:
        x = B.make_inputs(nbi, kind, 3).to(dev)
        mpc = BatchedMPC(pkm, H, 0.02, spec, integrator="euler", lr=lr, iters=it)
        pkm.set_option("latency_max_batch", 1 << 20)
        ms = timed(lambda: mpc.solve(x))
        print("%s euler H=%d it=%d B=%d via latency kernel: %.3f ms  %.0f solves/s" % (name, H, it, nbi, ms, nbi / ms * 1e3))
