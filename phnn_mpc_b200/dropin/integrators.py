"""Drop-in for src/integrators.py.  For models backed by the CUDA library a step or a whole
horizon is ONE kernel launch (op phnn_mpc::rollout); the per-step Python loop of the reference
(src/integrators.py:176-187, 239-248) is gone.  Returned tensors live on the device of ``y0``.

``rollout_trajectory_differentiable`` is differentiable where the reference's is: when autograd is recording and the
model's parameters (or y0 / controls) require a gradient, it returns ``training.trainable_rollout`` -- the backward is
the fused rollout + adjoint kernel in its training mode (dL/dtheta, dL/dy0, dL/dU), not an autograd tape.  The MPC path
gets dJ/dU from the in-kernel adjoint.  Not carried over: ``compare_integrators`` (dead code in the reference: it raises
under its own torch.no_grad())."""
import torch

from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import pack_of


def _pack(model):
    if not torch.cuda.is_available():
        raise RuntimeError("phnn_mpc_b200 has no CPU fallback: a CUDA device is required")
    return pack_of(model)


def _roll(model, y0, controls, dt, integrator, energy_mode):
    pk = _pack(model)
    iid = ops.integrator_id(integrator)
    y = y0.detach().to(device=pk.device, dtype=torch.float32).reshape(-1, pk.n)
    U = controls.detach().to(device=pk.device, dtype=torch.float32).reshape(y.shape[0], -1, pk.m)
    traj, en = ops.rollout(pk.handle, y, U, float(dt), iid, energy_mode)
    return traj.to(y0.device), (en.to(y0.device) if energy_mode else None)


def _one_step(model, y, u, dt, integrator, energy_mode=0):
    traj, en = _roll(model, y.reshape(-1, y.shape[-1]), u.reshape(-1, 1, u.shape[-1]), dt, integrator, energy_mode)
    return traj[:, 1], (en[:, 0] if en is not None else None)


def euler_step(model, y, u, dt):
    """y + dt f(y,u)  (src/integrators.py:13-36)"""
    return _one_step(model, y, u, dt, "euler")[0]


def rk4_step(model, y, u, dt):
    """classical RK4 with u held over the step (src/integrators.py:39-84)"""
    return _one_step(model, y, u, dt, "rk4")[0]


def rk4_step_with_energy(model, y, u, dt):
    """(y_next, H(y))  (src/integrators.py:87-125)"""
    return _one_step(model, y, u, dt, "rk4", 2)


def rollout_trajectory(model, y0, controls, dt, integrator="rk4"):
    """(trajectory [B,T+1,n], energies [B,T+1] = H(y_0..y_T))  (src/integrators.py:128-189)"""
    return _roll(model, y0, controls, dt, integrator, 2)


def rollout_trajectory_differentiable(model, y0, controls, dt, integrator="rk4", return_energies=False):
    """trajectory [B,T+1,n]; with return_energies also the reference's energy list
    [H(y0), H(y0), H(y1), ..., H(y_{T-1})]  (src/integrators.py:192-258)"""
    needs_grad = torch.is_grad_enabled() and (y0.requires_grad or controls.requires_grad or
                                              any(p.requires_grad for p in model.parameters()))
    if needs_grad and not return_energies and _pack(model).h <= 128:
        from phnn_mpc_b200.training import trainable_rollout
        return trainable_rollout(model, y0, controls, dt, integrator)
    traj, en = _roll(model, y0, controls, dt, integrator, 1 if return_energies else 0)
    return (traj, en) if return_energies else traj
