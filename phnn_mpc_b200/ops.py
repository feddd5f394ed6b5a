"""torch.library custom ops (namespace ``phnn_mpc``) over the C ABI of libphnn_mpc.so.

Each op takes the packed-weights handle (an int, see packing.PackedModel) and CUDA float32
tensors; PyTorch only provides device memory and the stream.  Reference functions replaced:

  phnn_mpc::forward    pHNN.forward / pHNN_Canonical.forward   (src/pHNN.py:52-100, src/pHNN_canonical.py:172-273)
  phnn_mpc::vjp        the autograd double-backward of the above (src/pHNN.py:73)
  phnn_mpc::rollout    rollout_trajectory[_differentiable]      (src/integrators.py:128-258)
  phnn_mpc::cost_grad  rollout + compute_cost + backward         (src/mpc_controller.py:75-141,176-194)
  phnn_mpc::mpc_solve  the whole Adam loop                       (src/mpc_controller.py:143-209,
                                                                  src/mpc_controller_canonical.py:163-228)
"""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

INTEGRATORS = {"euler": _lib.PHNN_EULER, "rk4": _lib.PHNN_RK4}


def integrator_id(name):
    if name not in INTEGRATORS:
        raise ValueError(f"Unknown integrator: {name}")  # same message as src/integrators.py:172
    return INTEGRATORS[name]


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError("phnn_mpc ops need CUDA tensors (%s is on %s); there is no CPU fallback" % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32" % name)
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dims(pack):
    L = _lib.lib()
    k, n, m, h = (ctypes.c_int() for _ in range(4))
    _lib.check(L.phnn_pack_dims(ctypes.c_void_p(pack), k, n, m, h), "phnn_pack_dims")
    return k.value, n.value, m.value, h.value


class _Cost:
    """Host-side phnn_cost_desc built from small CPU tensors."""

    def __init__(self, Q, R, xt, has_ub, umin, umax, xmin, xmax, bw):
        self.keep = []

        def hp(t):
            if t is None:
                return None
            a = t.detach().to("cpu", torch.float32).contiguous()
            self.keep.append(a)
            return ctypes.cast(a.data_ptr(), ctypes.POINTER(ctypes.c_float))

        d = _lib.CostDesc()
        d.Q, d.R, d.x_target = hp(Q), hp(R), hp(xt)
        d.has_u_bounds = int(has_ub)
        d.u_min, d.u_max = float(umin), float(umax)
        d.x_min, d.x_max = hp(xmin), hp(xmax)
        d.barrier_weight = float(bw)
        self.desc = d


@torch.library.custom_op("phnn_mpc::forward", mutates_args=())
def forward(pack: int, x: torch.Tensor, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    x, u = _chk(x, "x"), _chk(u, "u")
    B = x.shape[0]
    dx = torch.empty_like(x)
    H = torch.empty((B,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().phnn_forward(ctypes.c_void_p(pack), _p(x), _p(u), _p(dx), _p(H), B, _stream(x)),
                   "phnn_forward")
    return dx, H


@forward.register_fake
def _(pack, x, u):
    return torch.empty_like(x), x.new_empty((x.shape[0],))


@torch.library.custom_op("phnn_mpc::vjp", mutates_args=())
def vjp(pack: int, x: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    x, u, v = _chk(x, "x"), _chk(u, "u"), _chk(v, "v")
    B = x.shape[0]
    xb = torch.empty_like(x)
    ub = torch.empty_like(u)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().phnn_vjp(ctypes.c_void_p(pack), _p(x), _p(u), _p(v), _p(xb), _p(ub), B, _stream(x)),
                   "phnn_vjp")
    return xb, ub


@vjp.register_fake
def _(pack, x, u, v):
    return torch.empty_like(x), torch.empty_like(u)


def _forward_setup(ctx, inputs, output):
    pack, x, u = inputs
    ctx.pack = pack
    ctx.save_for_backward(x, u)
    # H carries no autograd formula here (the weights are not op inputs and dH/dx is not exported): differentiating
    # through it raises instead of silently contributing nothing
    ctx.mark_non_differentiable(output[1])


def _forward_backward(ctx, grad_dx, grad_H):
    # First-order in (x, u) through dx only; parameters receive no gradient from this op.
    x, u = ctx.saved_tensors
    xb, ub = vjp(ctx.pack, x, u, grad_dx.contiguous())
    return None, xb, ub


forward.register_autograd(_forward_backward, setup_context=_forward_setup)


@torch.library.custom_op("phnn_mpc::rollout", mutates_args=())
def rollout(pack: int, x0: torch.Tensor, U: torch.Tensor, dt: float, integrator: int,
            energy_mode: int) -> Tuple[torch.Tensor, torch.Tensor]:
    x0, U = _chk(x0, "x0"), _chk(U, "U")
    B, n = x0.shape
    T = U.shape[1]
    traj = torch.empty((B, T + 1, n), dtype=torch.float32, device=x0.device)
    en = torch.empty((B, T + 1) if energy_mode else (0,), dtype=torch.float32, device=x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().phnn_rollout(ctypes.c_void_p(pack), _p(x0), _p(U), _p(traj),
                                           _p(en) if energy_mode else None, B, T, float(dt), integrator, energy_mode,
                                           _stream(x0)), "phnn_rollout")
    return traj, en


@rollout.register_fake
def _(pack, x0, U, dt, integrator, energy_mode):
    B, n = x0.shape
    T = U.shape[1]
    return x0.new_empty((B, T + 1, n)), x0.new_empty((B, T + 1) if energy_mode else (0,))


def _workspace(pack, B, T, integrator, device):
    nbytes = _lib.lib().phnn_workspace_bytes(ctypes.c_void_p(pack), B, T, integrator)
    return torch.empty((max(nbytes, 4) + 3) // 4, dtype=torch.float32, device=device), nbytes


@torch.library.custom_op("phnn_mpc::cost_grad", mutates_args=())
def cost_grad(pack: int, x0: torch.Tensor, U: torch.Tensor, dt: float, integrator: int, Q: torch.Tensor,
              R: torch.Tensor, x_target: torch.Tensor, has_u_bounds: bool, u_min: float, u_max: float,
              x_min: Optional[torch.Tensor], x_max: Optional[torch.Tensor], barrier_weight: float,
              want_grad: bool, want_traj: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    x0, U = _chk(x0, "x0"), _chk(U, "U")
    B, n = x0.shape
    T = U.shape[1]
    cd = _Cost(Q, R, x_target, has_u_bounds, u_min, u_max, x_min, x_max, barrier_weight)
    cost = torch.empty((B,), dtype=torch.float32, device=x0.device)
    g = torch.empty_like(U) if want_grad else torch.empty((0,), dtype=torch.float32, device=x0.device)
    traj = torch.empty((B, T + 1, n) if want_traj else (0,), dtype=torch.float32, device=x0.device)
    ws, nbytes = _workspace(pack, B, T, integrator, x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().phnn_cost_grad(ctypes.c_void_p(pack), ctypes.byref(cd.desc), _p(x0), _p(U), _p(cost),
                                             _p(g) if want_grad else None, _p(traj) if want_traj else None, B, T,
                                             float(dt), integrator, _p(ws), nbytes, _stream(x0)), "phnn_cost_grad")
    return cost, g, traj


@cost_grad.register_fake
def _(pack, x0, U, dt, integrator, Q, R, x_target, has_u_bounds, u_min, u_max, x_min, x_max, barrier_weight,
      want_grad, want_traj):
    B, n = x0.shape
    T = U.shape[1]
    return (x0.new_empty((B,)), torch.empty_like(U) if want_grad else x0.new_empty((0,)),
            x0.new_empty((B, T + 1, n) if want_traj else (0,)))


@torch.library.custom_op("phnn_mpc::mpc_solve", mutates_args=())
def mpc_solve(pack: int, x0: torch.Tensor, U0: torch.Tensor, dt: float, integrator: int, Q: torch.Tensor,
              R: torch.Tensor, x_target: torch.Tensor, has_u_bounds: bool, u_min: float, u_max: float,
              x_min: Optional[torch.Tensor], x_max: Optional[torch.Tensor], barrier_weight: float, lr: float,
              beta1: float, beta2: float, eps: float, iters: int, return_mode: int,
              want_hist: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (U [B,T,m], cost_hist [iters,B] or empty, best_cost [B])."""
    x0, U0 = _chk(x0, "x0"), _chk(U0, "U0")
    B, n = x0.shape
    T = U0.shape[1]
    cd = _Cost(Q, R, x_target, has_u_bounds, u_min, u_max, x_min, x_max, barrier_weight)
    U = U0.clone()
    hist = torch.empty((iters, B) if want_hist else (0,), dtype=torch.float32, device=x0.device)
    best = torch.empty((B,), dtype=torch.float32, device=x0.device)
    ws, nbytes = _workspace(pack, B, T, integrator, x0.device)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().phnn_mpc_solve(ctypes.c_void_p(pack), ctypes.byref(cd.desc), _p(x0), _p(U),
                                             _p(hist) if want_hist else None, _p(best), B, T, float(dt), integrator,
                                             float(lr), float(beta1), float(beta2), float(eps), int(iters),
                                             int(return_mode), _p(ws), nbytes, _stream(x0)), "phnn_mpc_solve")
    return U, hist, best


@mpc_solve.register_fake
def _(pack, x0, U0, dt, integrator, Q, R, x_target, has_u_bounds, u_min, u_max, x_min, x_max, barrier_weight, lr,
      beta1, beta2, eps, iters, return_mode, want_hist):
    B = x0.shape[0]
    return torch.empty_like(U0), x0.new_empty((iters, B) if want_hist else (0,)), x0.new_empty((B,))
