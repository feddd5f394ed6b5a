"""Parity errors and timings of the tcgen05 kernel variants side by side (run on the GPU box).

    python tools/gpu_mode_check.py [modes, default 2,4] [--full]

Prints, per tensor_mode: forward / cost / dJdU / solve errors against the golden vectors recorded from the reference
and against the CPU oracle, then the time of the profiling-sized cfg4 slice (and of the full cfg4 job with --full)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import torch

import bench
from conftest import load_golden, rel_err
from phnn_mpc_b200 import ops
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.packing import PackedModel
from oracle.phnn_oracle import OracleModel, set_threads

modes = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else ["2", "4"])]
full = "--full" in sys.argv
set_threads(os.cpu_count() or 1)


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for name, kind in (("cartpole_h256", "phnn"), ("cartpole_h128", "phnn"), ("canonical", "canonical")):
    z, sd = load_golden(name)
    M = OracleModel(sd, kind)
    for mode in modes:
        try:
            pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
        except RuntimeError as ex:   # single-shape experiment builds
            print("%-14s skipped (%s)" % (name, str(ex)[:60]))
            break
        pk.set_option("tensor_min_batch", 0)
        pk.set_option("latency_max_batch", 0)
        pk.set_option("tensor_mode", mode)
        dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
        torch.cuda.synchronize()
        e_dx, e_H = rel_err(dx.cpu().numpy(), z["rand_dx"]), rel_err(H.cpu().numpy(), z["rand_H"])
        lo, hi = [float(v) for v in z["mpc_bounds"]]
        ca = (torch.from_numpy(z["mpc_Q"]), torch.from_numpy(z["mpc_R"]), torch.from_numpy(z["mpc_xt"]), True, lo, hi, None, None, 1000.0)
        dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
        cost, g, tr = ops.cost_grad(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, 1, *ca, True, True)
        torch.cuda.synchronize()
        e_c, e_g = rel_err(cost.cpu().numpy(), z["mpc_rk4_hist"][0]), rel_err(g.cpu().numpy(), z["mpc_rk4_grad0"])
        iters = z["mpc_rk4_hist"].shape[0]
        U, hist, best = ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, 1, *ca, lr, 0.9, 0.999, 1e-8, iters, 0, True)
        torch.cuda.synchronize()
        e_h, e_U = rel_err(hist.cpu().numpy(), z["mpc_rk4_hist"]), float(np.abs(U.cpu().numpy() - z["mpc_rk4_U_last"]).max())
        print("%-14s mode %d: fwd dx %.2e H %.2e | cost %.2e dJdU %.2e | solve hist %.2e U %.2e (golden, RK4, %d it)" % (
            name, mode, e_dx, e_H, e_c, e_g, e_h, e_U, iters), flush=True)
        if name == "cartpole_h256":
            # wide states (saturated units) and a 300-instance oracle comparison at H = 50
            dxw, Hw = ops.forward(pk.handle, cu(z["wide_x"]), cu(z["rand_u"][:16]))
            print("   wide states: dx %.2e H %.2e" % (rel_err(dxw.cpu().numpy(), z["wide_dx"]), rel_err(Hw.cpu().numpy(), z["wide_H"])))
            B, Hh = 300, 50
            x0 = bench.make_inputs(B, "phnn", 7).numpy()
            gq = torch.Generator().manual_seed(3)
            U0 = ((torch.rand(B, Hh, 1, generator=gq) * 2 - 1) * 3).numpy()
            Q = np.diag([10.0, 200.0, 1.0, 10.0]).astype(np.float32)
            Rm = np.array([[0.01]], np.float32)
            C = M.cost_struct(Q, Rm, np.zeros(4), -15.0, 15.0)
            cb = (torch.from_numpy(Q), torch.from_numpy(Rm), torch.zeros(4), True, -15.0, 15.0, None, None, 1000.0)
            Jo, go = M.cost_grad(C, x0, U0, 0.02, "rk4")
            cc, gg, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), 0.02, 1, *cb, True, False)
            Uo, ho, _ = M.mpc_solve(C, x0, U0, 0.02, "rk4", lr=0.015, iters=4)
            Ug, hg, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *cb, 0.015, 0.9, 0.999, 1e-8, 4, 0, True)
            torch.cuda.synchronize()
            print("   300 x H=50 vs oracle: cost %.2e dJdU %.2e | 4-it solve hist %.2e U %.2e" % (
                rel_err(cc.cpu().numpy(), Jo), rel_err(gg.cpu().numpy(), go), rel_err(hg.cpu().numpy(), ho),
                float(np.abs(Ug.cpu().numpy() - Uo).max())), flush=True)

# timings
sd = bench.load_fixture("cartpole_h256")
c = bench.cost_for("phnn")
spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
for wl in (["small", "cfg4_rk4"] if full else ["small"]):
    fixture, kind, B, H, iters, integ, lr, scaling, desc = bench.WORKLOADS[wl]
    x0 = bench.make_inputs(B, kind, 7).cuda()
    outs = {}
    for mode in modes:
        pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
        pk.set_option("tensor_mode", mode)
        mpc = BatchedMPC(pk, H, 0.02, spec, integrator=integ, lr=lr, iters=iters, return_mode="last")
        ms = timed(lambda: mpc.solve(x0))
        outs[mode] = mpc.solve(x0)["U"]
        print("%-9s mode %d: %.2f ms  (%.1f k solves/s)" % (wl, mode, ms, B / ms), flush=True)
    if len(modes) > 1:
        a, b = outs[modes[0]], outs[modes[1]]
        print("   max |U(mode %d) - U(mode %d)| = %.2e" % (modes[0], modes[1], float((a - b).abs().max())))
