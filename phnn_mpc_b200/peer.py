"""Fused result exchange of the multi-GPU solve.

Instances are sharded across ranks with no data-path collective; the only exchange is the gather of U* / best cost
(SURVEY.md section 8e).  ``PeerGather`` replaces the final NCCL all_gather: every rank allocates result buffers
``U_all [B_total, T, 1]`` and ``cost_all [B_total]`` through the C ABI (cudaMalloc + CUDA IPC handle), the handles
are exchanged once with ``torch.distributed``, and the solve kernel stores every finished 128-instance tile into the
buffers of ALL ranks over NVLink while the remaining tiles are still being solved (``phnn_mpc_solve_peer``).
``finish()`` is the only synchronisation: a 4-byte all_reduce on the solve's stream, after which every rank's buffers
hold the whole job.  Two buffer sets alternate, so a rank that runs ahead into the next solve never overwrites results
another rank is still reading (the all_reduce keeps ranks within one solve of each other).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


class _Buf:
    def __init__(self, L, nbytes, dev_index, group):
        self.L, self.dev = L, dev_index
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.phnn_peer_alloc(nbytes, dev_index, ctypes.byref(ptr), handle), "phnn_peer_alloc")
        self.local = ptr.value
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self.opened = [], []
        for r, h in enumerate(handles):
            if r == rank:
                self.ptrs.append(self.local)
                continue
            p = ctypes.c_void_p()
            _lib.check(L.phnn_peer_open(h, dev_index, ctypes.byref(p)), "phnn_peer_open")
            self.ptrs.append(p.value)
            self.opened.append(p.value)

    def close(self):
        for p in self.opened:
            self.L.phnn_peer_close(ctypes.c_void_p(p), self.dev)
        self.opened = []
        if self.local:
            self.L.phnn_peer_free(ctypes.c_void_p(self.local), self.dev)
            self.local = None


def _as_tensor(ptr, shape, device):
    """zero-copy float32 view of a raw device allocation (owned by the _Buf, which outlives the view)"""
    n = 1
    for s in shape:
        n *= s

    class _Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    with torch.cuda.device(device):
        return torch.as_tensor(_Holder(), device=device).view(*shape)


class PeerGather:
    """result buffers of one sharded job: B instances per rank, horizon T, `world` ranks on one node"""

    def __init__(self, B, T, world, rank, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGather needs an initialised torch.distributed process group")
        if world > 8:
            raise RuntimeError("PeerGather: at most 8 ranks (one NVSwitch node)")
        self.B, self.T, self.world, self.rank, self.group = int(B), int(T), int(world), int(rank), group
        self.device = torch.device(device)
        L = _lib.lib()
        tot = self.B * self.world
        self.sets = []
        for _ in range(2):
            bu = _Buf(L, tot * self.T * 4, self.device.index, group)
            bc = _Buf(L, tot * 4, self.device.index, group)
            self.sets.append((bu, bc, _as_tensor(bu.local, (tot, self.T, 1), self.device), _as_tensor(bc.local, (tot,), self.device)))
        self.cur = 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        dist.barrier(group=group)

    def desc(self):
        """phnn_peer_desc of the buffer set the next solve writes"""
        bu, bc, _, _ = self.sets[self.cur]
        d = _lib.PeerDesc()
        d.n = self.world
        d.offset = self.rank * self.B
        for r in range(self.world):
            d.U[r] = bu.ptrs[r]
            d.cost[r] = bc.ptrs[r]
        return d

    def finish(self):
        """stream-ordered barrier across ranks: afterwards U_all / cost_all of the solve just launched are complete on
        this rank; flips to the other buffer set for the next solve"""
        dist.all_reduce(self._flag, group=self.group)
        _, _, self.U_all, self.cost_all = self.sets[self.cur]
        self.cur ^= 1
        return self.U_all, self.cost_all

    def exchange_only(self, U, cost):
        """the exchange alone (for timing): copy this rank's finished shard into every rank's buffers, then the barrier"""
        bu, bc, _, _ = self.sets[self.cur]
        lo = self.rank * self.B
        for r in range(self.world):
            _as_tensor(bu.ptrs[r], (self.B * self.world, self.T, 1), self.device)[lo:lo + self.B].copy_(U, non_blocking=True)
            _as_tensor(bc.ptrs[r], (self.B * self.world,), self.device)[lo:lo + self.B].copy_(cost, non_blocking=True)
        return self.finish()

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for bu, bc, _, _ in self.sets:
            bu.close()
            bc.close()
        self.sets = []
