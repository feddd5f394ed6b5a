"""Fused peer-store exchange vs NCCL all_gather on N GPUs of one node (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/gpu_peer_check.py

Each rank solves its shard of a cfg4-shaped job twice: once with the solve kernel storing finished tiles straight into
every rank's result buffers (phnn_mpc_solve_peer + a 4-byte all_reduce), once followed by dist.all_gather.  The two
results must be bit-identical on every rank; prints the time of both exchange styles."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch
import torch.distributed as dist

import bench
from phnn_mpc_b200.batched import BatchedMPC, CostSpec
from phnn_mpc_b200.distributed import gather_shards
from phnn_mpc_b200.packing import PackedModel
from phnn_mpc_b200.peer import PeerGather

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
B, H, iters = int(os.environ.get("PEER_B", "9472")), 50, 3
sd = bench.load_fixture("cartpole_h256")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn", device=dev)
c = bench.cost_for("phnn")
spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
mpc = BatchedMPC(pk, H, 0.02, spec, integrator="rk4", lr=0.015, iters=iters, return_mode="last")
x0 = bench.make_inputs(B, "phnn", 7 + rank).to(dev)
peer = PeerGather(B, H, world, rank, dev)
ok = True
for rep in range(3):
    o = mpc.solve(x0, peer=peer)
    U_all, c_all = peer.finish()
    torch.cuda.synchronize()
    o2 = mpc.solve(x0)
    Ug = gather_shards(o2["U"], B * world)
    cg = gather_shards(o2["best_cost"], B * world)
    torch.cuda.synchronize()
    same = bool(torch.equal(U_all, Ug)) and bool(torch.equal(c_all, cg)) and bool(torch.equal(o["U"], o2["U"]))
    ok = ok and same
    print("rank %d rep %d: peer-store result == all_gather result: %s" % (rank, rep, same), flush=True)


def timed(fn, n=7):
    ts = []
    for i in range(n + 2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def with_peer():
    mpc.solve(x0, peer=peer); peer.finish()


def with_nccl():
    gather_shards(mpc.solve(x0)["U"], B * world)


t_solve = timed(lambda: mpc.solve(x0))
t_peer, t_nccl = timed(with_peer), timed(with_nccl)
print("rank %d: solve alone %.3f ms | solve + fused peer-store exchange %.3f ms | solve + NCCL all_gather %.3f ms" % (rank, t_solve, t_peer, t_nccl), flush=True)
peer.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
