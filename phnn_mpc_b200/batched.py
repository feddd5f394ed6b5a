"""Batched front-end over the phnn_mpc ops: B independent MPC instances per call.

The reference solves one instance per Python call (src/mpc_controller.py:143-209,
src/mpc_controller_canonical.py:163-273); instances are independent, so B of them are one
launch here.  Host tensors / NumPy arrays are staged to the GPU and results come back in the
caller's form; the arithmetic always runs in the CUDA kernel.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops
from .packing import PackedModel, pack_of


def _t(a, shape=None):
    if a is None:
        return None
    t = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).detach().to(torch.float32).cpu()
    return t.reshape(shape) if shape is not None else t


@dataclass
class CostSpec:
    """sum_{t<=H} (x_t-x*)^T Q (x_t-x*) + sum_{t<H} u_t^T R u_t (+ soft state bounds)."""
    Q: torch.Tensor
    R: torch.Tensor
    x_target: torch.Tensor
    u_min: Optional[float] = None
    u_max: Optional[float] = None
    x_min: Optional[torch.Tensor] = None
    x_max: Optional[torch.Tensor] = None
    barrier_weight: float = 1000.0

    @classmethod
    def make(cls, n, m, Q, R, x_target=None, u_min=None, u_max=None, x_min=None, x_max=None, barrier_weight=1000.0):
        Q = _t(Q)
        Q = torch.diag(Q) if Q.ndim == 1 else Q.reshape(n, n)
        R = _t(R)
        R = R.reshape(1, 1) * torch.eye(m) if R.ndim == 0 else (torch.diag(R) if R.ndim == 1 else R.reshape(m, m))
        xt = torch.zeros(n) if x_target is None else _t(x_target, (n,))
        return cls(Q.contiguous(), R.contiguous(), xt, u_min, u_max, _t(x_min, (n,)) if x_min is not None else None,
                   _t(x_max, (n,)) if x_max is not None else None, float(barrier_weight))

    def op_args(self):
        has = self.u_min is not None and self.u_max is not None  # both needed (src/mpc_controller.py:180)
        return (self.Q, self.R, self.x_target, bool(has), float(self.u_min) if has else 0.0,
                float(self.u_max) if has else 0.0, self.x_min, self.x_max, float(self.barrier_weight))


def as_pack(model, device=None):
    if isinstance(model, PackedModel):
        return model
    return pack_of(model, device)


def _dev(x, device):
    t = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def forward(model, x, u):
    pk = as_pack(model)
    x = _dev(x, pk.device).reshape(-1, pk.n)
    u = _dev(u, pk.device).reshape(-1, pk.m)
    return ops.forward(pk.handle, x, u)


def rollout(model, x0, U, dt, integrator="rk4", energy_mode=0):
    """traj [B,T+1,n] (and energies [B,T+1] for energy_mode 1|2) on the pack's device."""
    pk = as_pack(model)
    x0 = _dev(x0, pk.device).reshape(-1, pk.n)
    U = _dev(U, pk.device).reshape(x0.shape[0], -1, pk.m)
    traj, en = ops.rollout(pk.handle, x0, U, float(dt), ops.integrator_id(integrator), int(energy_mode))
    return (traj, en) if energy_mode else traj


def _solve_peer_call(pk, x0, U0, dt, integ, cost_args, lr, b1, b2, eps, iters, rmode, want_hist, peer):
    """phnn_mpc_solve_peer through ctypes (the torch.library op has no argument for the peer table)"""
    import ctypes
    from . import _lib
    B, T = x0.shape[0], U0.shape[1]
    cd = ops._Cost(*cost_args)
    U = U0.clone()
    hist = torch.empty((iters, B) if want_hist else (0,), dtype=torch.float32, device=x0.device)
    best = torch.empty((B,), dtype=torch.float32, device=x0.device)
    ws, nbytes = ops._workspace(pk.handle, B, T, integ, x0.device)
    d = peer.desc()
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().phnn_mpc_solve_peer(ctypes.c_void_p(pk.handle), ctypes.byref(cd.desc), ops._p(x0), ops._p(U),
                                                  ops._p(hist) if want_hist else None, ops._p(best), B, T, float(dt), integ,
                                                  float(lr), float(b1), float(b2), float(eps), int(iters), int(rmode),
                                                  ops._p(ws), nbytes, ctypes.byref(d), ops._stream(x0)), "phnn_mpc_solve_peer")
    return U, hist, best


class BatchedMPC:
    """B-instance gradient MPC: iters x {clamp, rollout, cost, adjoint, Adam} in one launch."""

    def __init__(self, model, horizon, dt, cost: CostSpec, integrator="euler", lr=0.1, iters=50, betas=(0.9, 0.999),
                 eps=1e-8, return_mode="last", device=None):
        self.model = model
        self.device = device
        self.horizon, self.dt, self.cost = int(horizon), float(dt), cost
        self.integrator = integrator
        self.integrator_id = ops.integrator_id(integrator)
        self.lr, self.iters, self.betas, self.eps = float(lr), int(iters), betas, float(eps)
        if return_mode not in ("last", "best"):
            raise ValueError("return_mode must be 'last' or 'best'")
        self.return_mode = return_mode

    @property
    def pack(self):
        return as_pack(self.model, self.device)

    def cost_and_grad(self, x0, U, want_grad=True, want_traj=False):
        pk = self.pack
        x0 = _dev(x0, pk.device).reshape(-1, pk.n)
        U = _dev(U, pk.device).reshape(x0.shape[0], self.horizon, pk.m)
        cost, g, traj = ops.cost_grad(pk.handle, x0, U, self.dt, self.integrator_id, *self.cost.op_args(),
                                      bool(want_grad), bool(want_traj))
        return cost, (g if want_grad else None), (traj if want_traj else None)

    def solve(self, x0, U0=None, want_hist=False, peer=None):
        """x0 [B,n]; U0 [B,H,m] or None (zeros, the controllers' cold start).
        Returns dict(U [B,H,m], u0 [B,m], best_cost [B], cost_hist [iters,B]|None) on the device.
        peer: a peer.PeerGather -- the kernel also stores the results into every rank's result buffers (multi-GPU)."""
        pk = self.pack
        x0 = _dev(x0, pk.device).reshape(-1, pk.n)
        B = x0.shape[0]
        if U0 is None:
            U0 = torch.zeros((B, self.horizon, pk.m), dtype=torch.float32, device=pk.device)
        else:
            U0 = _dev(U0, pk.device).reshape(B, self.horizon, pk.m)
        if peer is not None:
            return self._solve_peer(pk, x0, U0, want_hist, peer)
        U, hist, best = ops.mpc_solve(pk.handle, x0, U0, self.dt, self.integrator_id, *self.cost.op_args(), self.lr,
                                      self.betas[0], self.betas[1], self.eps, self.iters,
                                      0 if self.return_mode == "last" else 1, bool(want_hist))
        return {"U": U, "u0": U[:, 0], "best_cost": best, "cost_hist": hist if want_hist else None}

    def _solve_peer(self, pk, x0, U0, want_hist, peer):
        U, hist, best = _solve_peer_call(pk, ops._chk(x0, "x0"), ops._chk(U0, "U0"), self.dt, self.integrator_id,
                                         self.cost.op_args(), self.lr, self.betas[0], self.betas[1], self.eps, self.iters,
                                         0 if self.return_mode == "last" else 1, bool(want_hist), peer)
        return {"U": U, "u0": U[:, 0], "best_cost": best, "cost_hist": hist if want_hist else None}
