"""Multi-GPU: MPC instances are independent, so the batch is cut into contiguous shards, one
process per GPU (torch.distributed); the only collective is the final gather of the results
(SURVEY.md section 8e).  Works with the nccl backend on GPUs and with gloo for host-side tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(B, world_size, rank):
    """Contiguous shard [lo, hi) of B instances for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local, B, group=None, dim=0):
    """all_gather of per-rank shards (instances along `dim`) into the full tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    if dim != 0:
        return gather_shards(local.movedim(dim, 0).contiguous(), B, group).movedim(0, dim)
    ws = dist.get_world_size(group)
    sizes = [shard_bounds(B, ws, r) for r in range(ws)]
    maxn = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == maxn for lo, hi in sizes):
        # equal shards: one all_gather_into_tensor straight into the result (no padding, no concatenation)
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)


# results of a solve and the axis their instances run along; every rank gathers exactly these keys, in this order
# (collective participation must never depend on a per-rank shape test: shard sizes differ by one across ranks)
GATHER_KEYS = (("U", 0), ("u0", 0), ("best_cost", 0), ("cost_hist", 1))


def sharded_solve(solve_fn, x0, U0=None, group=None):
    """Run `solve_fn(x0_shard, U0_shard) -> dict of tensors with leading shard dim` on this
    rank's contiguous shard and gather U / u0 / best_cost (instances along dim 0) and cost_hist ([iters, B]:
    instances along dim 1) on every rank.  Which keys are gathered depends only on which are present (not None) in
    the result, which is the same on every rank."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x0.shape[0]
    lo, hi = shard_bounds(B, ws, rank)
    out = solve_fn(x0[lo:hi], None if U0 is None else U0[lo:hi])
    unknown = [k for k, v in out.items() if v is not None and k not in dict(GATHER_KEYS)]
    if unknown:
        raise KeyError("sharded_solve: no gather rule for result(s) %s (known: %s)" % (unknown, [k for k, _ in GATHER_KEYS]))
    return {k: gather_shards(out[k], B, group, dim) for k, dim in GATHER_KEYS if out.get(k) is not None}
