"""Throughput of the device-resident batched closed loop: plants x sim-steps per second."""
import os, sys, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import load_golden
from phnn_mpc_b200.packing import PackedModel
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.closed_loop import ClosedLoopBatch
z, sd = load_golden("cartpole_h128")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
spec = CostSpec.make(4, 1, [10.0, 200.0, 1.0, 10.0], [0.01], None, -15.0, 15.0)
mpc = BatchedMPC(pk, 20, 0.02, spec, integrator="euler", lr=0.015, iters=30)
loop = ClosedLoopBatch(mpc, warm_start=False, tolerance=[0.1, 0.1, 0.05, 0.05], min_duration=0.2)
for B, steps in ((1, 50), (16384, 25)):
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1, 1, size=(B, 4)) * [0.5, 0.1, 0.2, 0.2]
    loop.run(x0, 2); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = loop.run(x0, steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"closed_loop": "cfg1 controller (h=128, H=20, 30 it, Euler) + plant, B=%d plants x %d steps" % (B, steps),
                      "s": dt, "plant_steps_per_s": B * steps / dt, "ms_per_sim_step": dt / steps * 1e3,
                      "alive": int((out["done_step"] < 0).sum().item()), "stable": int(out["stability_achieved"].sum().item())}))
