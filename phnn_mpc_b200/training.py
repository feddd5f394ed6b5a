"""Weight-gradient (training) mode: a rollout that autograd can differentiate with respect to the model's PARAMETERS.

The reference trains by unrolling ``model(x_t, u_t)`` over a data sequence and calling ``loss.backward()``; autograd
then runs a double backward through every ``torch.autograd.grad`` inside ``forward`` (scripts/train_cartpole_phnn.py:
108-178, scripts/train_cartpole_phnn_canonical.py:83-196).  Here the whole unrolled trajectory is ONE differentiable op:

    traj = trainable_rollout(model, x0, U, dt, "euler")      # [B, T+1, n], CUDA kernel (phnn_rollout)
    loss = any torch expression of traj                       # position / angle / velocity terms of the scripts
    loss.backward()                                           # phnn_rollout_vjp: dL/dx0, dL/dU and dL/dtheta

The backward is the fused rollout + discrete adjoint kernel in its training mode plus the contraction kernel
(csrc: MODE_PARAMGRAD, atb_kernel); no autograd tape of the model exists.  Parameters that the reference itself treats
as constants on this path get no gradient either: ``G_fixed`` / ``G`` / the canonical ``J`` (buffers) and the mass-matrix
scalars (the reference reads them with ``.item()``, src/mass_matrix.py:296-306).  ``H_net``'s output bias does not enter the
dynamics.
"""
import ctypes

import torch

from . import _lib, ops
from .packing import pack_of

# state_dict key of each gradient slot of phnn_param_grads
_SLOTS = {"W1": "H_net.net.0.weight", "b1": "H_net.net.0.bias", "W2": "H_net.net.2.weight", "b2": "H_net.net.2.bias",
          "W3": "H_net.net.4.weight", "Wr1": "R_net.net.0.weight", "br1": "R_net.net.0.bias", "Wr2": "R_net.net.2.weight",
          "br2": "R_net.net.2.bias", "Wg1": "G_net.net.0.weight", "bg1": "G_net.net.0.bias", "Wg2": "G_net.net.2.weight",
          "bg2": "G_net.net.2.bias", "J": "J"}


def rollout_vjp(pk, x0, U, gtraj, dt, integrator, want):
    """dL/dx0 [B,n], dL/dU [B,T,m] and {state_dict key: gradient} for the keys in `want`, from gtraj = dL/dtraj."""
    L = _lib.lib()
    x0, U, gtraj = ops._chk(x0, "x0"), ops._chk(U, "U"), ops._chk(gtraj, "gtraj")
    B, n = x0.shape
    T = U.shape[1]
    integ = ops.integrator_id(integrator)
    dev = x0.device
    dx0 = torch.empty_like(x0)
    dU = torch.empty_like(U)
    g = _lib.ParamGrads()
    out = {}
    for slot, key in _SLOTS.items():
        if key in want:
            out[key] = torch.empty(want[key], dtype=torch.float32, device=dev)
            setattr(g, slot, out[key].data_ptr())
    rd = None
    if "R_diag_raw" in want:
        rd = torch.empty(n, dtype=torch.float32, device=dev)
        g.r_diag = rd.data_ptr()
    nbytes = L.phnn_rollout_vjp_workspace_bytes(ctypes.c_void_p(pk.handle), B, T, integ)
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.phnn_rollout_vjp(ctypes.c_void_p(pk.handle), ops._p(x0), ops._p(U), ops._p(gtraj), ops._p(dx0), ops._p(dU),
                                      ctypes.byref(g), B, T, float(dt), integ, ops._p(ws), nbytes, ops._stream(x0)),
                   "phnn_rollout_vjp")
    if rd is not None:
        out["__r_diag"] = rd
    return dx0, dU, out


class _TrainableRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, dt, integrator, x0, U, *params):
        pk = pack_of(module)
        ctx.module, ctx.dt, ctx.integrator, ctx.pk = module, dt, integrator, pk
        ctx.in_device = x0.device
        xd = x0.detach().to(pk.device, torch.float32).contiguous()
        Ud = U.detach().to(pk.device, torch.float32).contiguous()
        traj, _ = ops.rollout(pk.handle, xd, Ud, float(dt), ops.integrator_id(integrator), 0)
        ctx.save_for_backward(xd, Ud)
        return traj.to(x0.device)

    @staticmethod
    def backward(ctx, gtraj):
        xd, Ud = ctx.saved_tensors
        named = [(k, p) for k, p in ctx.module.named_parameters()]
        if any(k == "M_net.L_tril" and p.requires_grad for k, p in named):
            # the reference keeps the constant mass matrix in the autograd graph (src/mass_matrix.py:130-147); this path has
            # no cotangent for it, and a silently missing gradient would be worse than an error
            raise NotImplementedError("training mode: no gradient for M_net.L_tril (constant mass matrix); freeze it "
                                      "(requires_grad_(False)) or train it with the reference")
        want = {k: tuple(p.shape) for k, p in named if p.requires_grad and (k in _SLOTS.values() or k == "R_diag_raw")}
        dx0, dU, gr = rollout_vjp(ctx.pk, xd, Ud, gtraj.to(ctx.pk.device, torch.float32).contiguous(), ctx.dt, ctx.integrator, want)
        grads = []
        for k, p in named:
            if k == "R_diag_raw" and "__r_diag" in gr:
                # r = softplus(R_diag_raw) + 1e-4 (src/pHNN_canonical.py:162): dr/draw = sigmoid(raw)
                grads.append((gr["__r_diag"].to(p.device) * torch.sigmoid(p.detach())).reshape(p.shape))
            elif k in gr:
                grads.append(gr[k].to(p.device).reshape(p.shape))
            else:
                grads.append(None)   # constants on this path (see module docstring)
        return (None, None, None, dx0.to(ctx.in_device), dU.to(ctx.in_device)) + tuple(grads)


def trainable_rollout(model, y0, controls, dt, integrator="euler"):
    """traj [B, T+1, n] of ``integrators.rollout_trajectory_differentiable`` (src/integrators.py:192-258), differentiable
    with respect to ``model.parameters()``, ``y0`` and ``controls``."""
    y0 = y0 if y0.ndim == 2 else y0.reshape(1, -1)
    if integrator not in ops.INTEGRATORS:
        raise ValueError(f"Unknown integrator: {integrator}")
    for net in (getattr(model, "H_net", None), getattr(model, "R_net", None), getattr(model, "G_net", None)):
        if net is not None and getattr(net, "dropout_p", 0.0) > 0.0:
            # the gradient slots are keyed by the dropout-free layer indices, and a training-mode forward would need the
            # reference's random masks: neither exists on this path
            raise NotImplementedError("training mode is built for MLPs without Dropout")
    params = [p for _, p in model.named_parameters()]
    return _TrainableRollout.apply(model, float(dt), integrator, y0, controls.reshape(y0.shape[0], -1, controls.shape[-1]), *params)
