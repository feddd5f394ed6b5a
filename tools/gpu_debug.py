"""Quick on-GPU numerics printout (debug aid; the judged checks are tests/test_gpu_parity.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden, rel_err
from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import PackedModel

KINDS = {"pendulum": "phnn", "cartpole_h128": "phnn", "cartpole_h256": "phnn", "canonical": "canonical"}
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
TC = os.environ.get("PHNN_TC")
names = [a for a in sys.argv[1:]] or list(KINDS)
for name in names:
    z, sd = load_golden(name)
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, KINDS[name])
    if TC is not None:
        pk.set_option("tensor_min_batch", 0)
        pk.set_option("tensor_mode", int(TC))
    print(name, "packed; tensor_mode", pk.get_option("tensor_mode"), flush=True)
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"])); torch.cuda.synchronize()
    print("  fwd  dx %.2e  H %.2e" % (rel_err(dx.cpu().numpy(), z["rand_dx"]), rel_err(H.cpu().numpy(), z["rand_H"])), flush=True)
    xb, ub = ops.vjp(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]), cu(z["rand_v"])); torch.cuda.synchronize()
    print("  vjp  xb %.2e  ub %.2e" % (rel_err(xb.cpu().numpy(), z["rand_gx"]), rel_err(ub.cpu().numpy(), z["rand_gu"])), flush=True)
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    ca = (torch.from_numpy(z["mpc_Q"]), torch.from_numpy(z["mpc_R"]), torch.from_numpy(z["mpc_xt"]), True, lo, hi, None, None, 1000.0)
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    for integ, iid in (("euler", 0), ("rk4", 1)):
        p = "mpc_%s_" % integ
        tr2, _ = ops.rollout(pk.handle, cu(z["mpc_x0"]), cu(np.clip(z["mpc_U0"], lo, hi)), dt, iid, 0); torch.cuda.synchronize()
        if p + "traj0" in z.files: print("  %s rollout %.2e" % (integ, rel_err(tr2.cpu().numpy(), z[p + "traj0"])), flush=True)
        cost, g, tr = ops.cost_grad(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, True, True); torch.cuda.synchronize()
        print("  %s cost %.2e grad %.2e" % (integ, rel_err(cost.cpu().numpy(), z[p + "hist"][0]), rel_err(g.cpu().numpy(), z[p + "grad0"])), flush=True)
        iters = z[p + "hist"].shape[0]
        for mode, key in ((0, "U_last"), (1, "U_best")):
            U, hist, best = ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, lr, 0.9, 0.999, 1e-8, iters, mode, True)
            torch.cuda.synchronize()
            print("  %s solve mode %d hist %.2e U abs %.2e best %.2e" % (integ, mode, rel_err(hist.cpu().numpy(), z[p + "hist"]),
                  np.abs(U.cpu().numpy() - z[p + key]).max(), rel_err(best.cpu().numpy(), z[p + "best"])), flush=True)
print("OK")
