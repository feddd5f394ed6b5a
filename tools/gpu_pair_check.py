"""CTA pairs (option "tensor_pair": tcgen05 cta_group::2, two tiles per cluster sharing each weight tile) against the
single-CTA launch of the same kernel: results must be bit-identical (same MMAs per row, same element code); then the
time of both on bench workloads.  Run on the GPU box:  python tools/gpu_pair_check.py [workload ...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.packing import PackedModel

def model(fixture, kind):
    sd = bench.load_fixture(fixture)
    c = bench.cost_for(kind)
    spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
    pk.set_option("tensor_min_batch", 0)
    pk.set_option("latency_max_batch", 0)
    pk.set_option("tensor_mode", 4)
    return pk, spec

for fixture, kind in (("cartpole_h256", "phnn"), ("cartpole_h128", "phnn"), ("canonical", "canonical")):
    try:
        pk, spec = model(fixture, kind)
    except Exception as ex:  # single-shape experiment builds
        print("%-14s skipped (%s)" % (fixture, str(ex)[:70]), flush=True)
        continue
    for B, H, iters in ((256, 6, 2), (512, 8, 3), (128 * 300, 5, 3), (128 * 22 - 5, 4, 2)):
        x0 = bench.make_inputs(B, kind, 11).cuda()
        mpc = BatchedMPC(pk, H, 0.02, spec, integrator="rk4", lr=0.015, iters=iters, return_mode="last")
        outs = []
        for pair in (0, 1):
            pk.set_option("tensor_pair", pair)
            o = mpc.solve(x0, want_hist=True)
            torch.cuda.synchronize()
            outs.append((o["U"].clone(), o["cost_hist"].clone()))
        dU = (outs[0][0] - outs[1][0]).abs().max().item()
        dh = (outs[0][1] - outs[1][1]).abs().max().item()
        print("%-14s B=%6d H=%d it=%d  pair vs single: max|dU| %.3e  max|dhist| %.3e  (|U| %.3e, hist %.6e)  %s" % (
            fixture, B, H, iters, dU, dh, outs[0][0].abs().max().item(), outs[0][1].double().sum().item(),
            "BIT-IDENTICAL" if dU == 0 and dh == 0 else "DIFFERENT"), flush=True)

sd = bench.load_fixture("cartpole_h256")
for wl in (sys.argv[1:] or ["small", "cfg4_rk4"]):
    fixture, kind, B, H, iters, integ, lr, scaling, desc = bench.WORKLOADS[wl]
    pk, spec = model(fixture, kind)
    x0 = bench.make_inputs(B, kind, 7).cuda()
    mpc = BatchedMPC(pk, H, 0.02, spec, integrator=integ, lr=lr, iters=iters, return_mode="last")
    res = {}
    for rnd in range(2):
        for pair in (0, 1):
            pk.set_option("tensor_pair", pair)
            for _ in range(1 if rnd else 2):
                mpc.solve(x0)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = mpc.solve(x0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            U = out["U"].double()
            print("%-9s pair=%d  %.2f ms (%.1f k solves/s)  checksum U %.9e" % (wl, pair, np.median(ts), B / np.median(ts), U.sum().item()), flush=True)
