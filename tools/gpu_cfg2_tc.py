"""Pendulum model (n = 2, h = 64, learned G) on the forward-only tcgen05 instantiation: parity against the golden
rollouts / the oracle, and timing of BASELINE cfg2 (4096 x H=100, RK4) per route plus a batch sweep."""
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
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
print("default options: tensor_mode %d tensor_fwd_min_batch %d latency_max_batch %d" % (pk.get_option("tensor_mode"),
      pk.get_option("tensor_fwd_min_batch"), pk.get_option("latency_max_batch")))
lat_default = pk.get_option("latency_max_batch")
def route(tc):
    pk.set_option("tensor_fwd_min_batch", 1 if tc else 0)
for tc in (0, 1):
    route(tc)
    name = "tcgen05" if tc else "latency"
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
    print("%s forward: dx %.2e H %.2e" % (name, rel_err(dx.cpu().numpy(), z["rand_dx"]), rel_err(H.cpu().numpy(), z["rand_H"])))
    for integ, iid in (("rk4", 1), ("euler", 0)):
        tr, en = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, iid, 1)
        tr, en = tr.cpu().numpy(), en.cpu().numpy()
        o64 = OracleModel(sd, "phnn", np.float64).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
        o32 = OracleModel(sd, "phnn", np.float32).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
        print("%s cfg2 %s: traj vs ref %.2e  en %.2e | vs f64 oracle %.2e (f32 oracle %.2e) | first 10 steps %.2e" % (
            name, integ, rel_err(tr, z["cfg2_traj_" + integ]), rel_err(en, z["cfg2_en_" + integ]), rel_err(tr, o64), rel_err(o32, o64),
            rel_err(tr[:, :11], z["cfg2_traj_" + integ][:, :11])))
        tr2, en2 = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, iid, 2)
        print("   energy ordering 2: traj %.2e" % rel_err(tr2.cpu().numpy(), tr))
# timing
g = torch.Generator().manual_seed(3)
def timeit(B, T=100):
    x0 = ((torch.rand(B, 2, generator=g) * 2 - 1) * torch.tensor([2.0, 2.0])).cuda()
    U = ((torch.rand(B, T, 1, generator=g) * 2 - 1)).cuda()
    for _ in range(3): ops.rollout(pk.handle, x0, U, 0.05, 1, 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.rollout(pk.handle, x0, U, 0.05, 1, 0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out[0]
for B in (int(b) for b in (os.environ.get("SWEEP") or "128,256,512,1024,2048,4096,18944,65536").split(",")):
    route(0); pk.set_option("latency_max_batch", 1 << 30)
    tl, a = timeit(B)
    pk.set_option("latency_max_batch", 0)
    tf, _ = timeit(B)
    route(1)
    pk.set_option("tensor_fwd_sparse", 0)
    td, bd = timeit(B)
    pk.set_option("tensor_fwd_sparse", 1)
    tt, b = timeit(B)
    pk.set_option("latency_max_batch", lat_default)
    print("B=%6d x H=100 RK4: latency %.3f ms  FP32-FMA %.3f ms  tcgen05 128-row tiles %.3f ms  tcgen05 default (64-row tiles up to 9472) %.3f ms  (%.1f M inst-steps/s)  dense vs sparse max|diff| %.2e" % (
          B, tl, tf, td, tt, B * 100 / tt / 1e3, (bd - b).abs().max().item()), flush=True)
