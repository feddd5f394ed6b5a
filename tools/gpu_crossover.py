"""Crossover of the three kernels (latency / FP32-FMA batched / tcgen05) vs batch size (debug aid)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import load_golden
from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import PackedModel
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for name, kind, H, iters, lr in (("cartpole_h128", "phnn", 20, 30, 0.015), ("canonical", "canonical", 10, 50, 0.03), ("pendulum", "phnn", 20, 10, 0.05)):
    z, sd = load_golden(name)
    n = 2 if name == "pendulum" else 4
    spec = CostSpec.make(n, 1, [10.0, 200.0, 1.0, 10.0][:n], [0.01], None, -15.0, 15.0)
    for B in [int(b) for b in (os.environ.get("SWEEP") or "148,296,592,1184,2368,4736,9472").split(",")]:
        x0 = (torch.rand(B, n) * 0.2 - 0.1).cuda()
        res = []
        for route in ("lat", "ffma", "tc"):
            pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
            if route == "lat":
                pk.set_option("latency_max_batch", 1 << 30)
            elif route == "ffma":
                pk.set_option("latency_max_batch", 0)
                if pk.get_option("tensor_mode") > 0: pk.set_option("tensor_mode", 0)
            else:
                if pk.get_option("tensor_mode") <= 0: res.append(float("nan")); continue
                pk.set_option("latency_max_batch", 0); pk.set_option("tensor_min_batch", 0)
            mpc = BatchedMPC(pk, H, 0.02, spec, integrator="euler", lr=lr, iters=iters)
            res.append(timed(lambda: mpc.solve(x0)))
        print("%-14s B=%5d  lat %8.2f ms  ffma %8.2f ms  tc %8.2f ms" % (name, B, *res), flush=True)
