"""Tiny run of all three kernels for compute-sanitizer (memcheck / racecheck)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import load_golden
from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import PackedModel
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for name, kind in (("cartpole_h128", "phnn"), ("canonical", "canonical"), ("pendulum", "phnn")):
    z, sd = load_golden(name)
    n = 2 if name == "pendulum" else 4
    for route in ("lat", "ffma", "tc"):
        if which not in ("all", route):
            continue
        pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
        if route == "tc" and pk.get_option("tensor_mode") <= 0:
            continue
        pk.set_option("latency_max_batch", 1 << 20 if route == "lat" else 0)
        if pk.get_option("tensor_mode") > 0:
            pk.set_option("tensor_mode", 3 if route == "tc" else 0)
        B = 40 if route != "tc" else 130
        x0 = (torch.rand(B, n) * 0.2 - 0.1).cuda()
        U0 = (torch.rand(B, 3, 1) * 2 - 1).cuda()
        Q = torch.eye(n); Rm = torch.tensor([[0.01]])
        ops.forward(pk.handle, x0, U0[:, 0].contiguous())
        ops.vjp(pk.handle, x0, U0[:, 0].contiguous(), x0)
        ops.rollout(pk.handle, x0, U0, 0.02, 1, 2)
        ops.cost_grad(pk.handle, x0, U0, 0.02, 1, Q, Rm, torch.zeros(n), True, -0.5, 0.5, None, None, 1000.0, True, True)
        ops.mpc_solve(pk.handle, x0, U0, 0.02, 0, Q, Rm, torch.zeros(n), True, -0.5, 0.5, None, None, 1000.0, 0.01, 0.9, 0.999, 1e-8, 2, 1, True)
        torch.cuda.synchronize()
        print(name, route, "ok", flush=True)
