"""A/B of the tanh variant on the long pendulum rollout (debug aid)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from conftest import load_golden, rel_err
from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import PackedModel
from oracle.phnn_oracle import OracleModel
z, sd = load_golden("pendulum")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()
M32, M64 = OracleModel(sd, "phnn"), OracleModel(sd, "phnn", np.float64)
for integ, iid in (("euler", 0), ("rk4", 1)):
    tr, _ = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, iid, 0)
    tr = tr.cpu().numpy()
    ref = z["cfg2_traj_" + integ]
    o32 = M32.rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    o64 = M64.rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    print(integ, "gpu-vs-ref %.2e  gpu-vs-f64 %.2e  ref-vs-f64 %.2e  o32-vs-f64 %.2e  o32-vs-ref %.2e" % (
        rel_err(tr, ref), rel_err(tr, o64), rel_err(ref, o64), rel_err(o32, o64), rel_err(o32, ref)))
    # per-instance worst
    e_gpu = np.abs(tr - o64).max(axis=(1, 2)); e_ref = np.abs(ref - o64).max(axis=(1, 2))
    print("   worst instance gpu %.2e (idx %d)  ref %.2e (idx %d)" % (e_gpu.max(), e_gpu.argmax(), e_ref.max(), e_ref.argmax()))
