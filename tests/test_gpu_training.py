"""Weight-gradient (training) mode, SURVEY.md section 8f row 3: dL/dtheta of the reference's training losses from
``training.trainable_rollout`` (the fused rollout + adjoint kernel in MODE_PARAMGRAD plus the contraction kernel) against
the gradients the REFERENCE's autograd produced for the same weights and batch (tests/golden/train_grads.npz, recorded by
make_golden.gen_train from scripts/train_cartpole_phnn.py:108-178 and scripts/train_cartpole_phnn_canonical.py:83-196).
Tolerance: 1e-4 of the largest entry of each gradient tensor (a 15-step unroll)."""
import os

import numpy as np
import pytest
import torch

from conftest import CONFIGS, GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _load():
    z = np.load(os.path.join(GOLDEN, "train_grads.npz"))
    return z, torch.from_numpy(z["x_batch"]), torch.from_numpy(z["u_batch"]), float(z["dt"])


def _model(kind, z):
    from phnn_mpc_b200.dropin.pHNN import pHNN
    from phnn_mpc_b200.dropin.pHNN_canonical import pHNN_Canonical
    cls, pre = (pHNN, "sd_phnn/") if kind == "phnn" else (pHNN_Canonical, "sd_canon/")
    m = cls(os.path.join(CONFIGS, "cartpole_phnn.yaml"))
    m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
    return m


def _check_param_grads(model, z, prefix, skip=()):
    worst = 0.0
    for k, p in model.named_parameters():
        ref = z[prefix + "/" + k]
        if k in skip:
            continue
        got = np.zeros_like(ref) if p.grad is None else p.grad.detach().cpu().numpy()
        scale = np.abs(ref).max()
        if scale == 0.0:                                     # parameters the reference gives no gradient on this path
            assert np.abs(got).max() == 0.0, k
            continue
        err = np.abs(got - ref).max() / scale
        worst = max(worst, err)
        assert err < TOL, (k, err)
    return worst


def test_phnn_euler_training_loss_gradients():
    """scripts/train_cartpole_phnn.py:108-160 (position MSE + angle cosine distance + velocity MSE over a 15-step Euler unroll)"""
    from phnn_mpc_b200.training import trainable_rollout
    z, xb, ub, dt = _load()
    model = _model("phnn", z)
    x0 = xb[:, 0, :].clone().requires_grad_(True)
    traj = trainable_rollout(model, x0, ub[:, :-1, :], dt, "euler")
    assert rel_err(traj.detach().numpy(), z["phnn_euler/traj"]) < TOL
    mse = torch.nn.MSELoss()
    loss = mse(traj[:, :, 0], xb[:, :, 0]) + torch.mean(1 - torch.cos(traj[:, :, 1] - xb[:, :, 1])) + mse(traj[:, :, 2:], xb[:, :, 2:])
    assert abs(loss.item() - float(z["phnn_euler/loss"])) < TOL * abs(float(z["phnn_euler/loss"]))
    loss.backward()
    _check_param_grads(model, z, "phnn_euler")
    assert rel_err(x0.grad.numpy(), z["phnn_euler/x0_grad"]) < TOL


def test_phnn_rk4_trajectory_matching_gradients_through_dropin_rollout():
    """the drop-in integrators.rollout_trajectory_differentiable is differentiable w.r.t. the parameters and the controls"""
    from phnn_mpc_b200.dropin.integrators import rollout_trajectory_differentiable
    z, xb, ub, dt = _load()
    model = _model("phnn", z)
    U = ub[:, :-1, :].clone().requires_grad_(True)
    traj = rollout_trajectory_differentiable(model, xb[:, 0, :].clone().requires_grad_(True), U, dt, "rk4")
    loss = ((traj - xb) ** 2).mean()
    assert abs(loss.item() - float(z["phnn_rk4/loss"])) < TOL * abs(float(z["phnn_rk4/loss"]))
    loss.backward()
    _check_param_grads(model, z, "phnn_rk4")
    assert rel_err(U.grad.numpy(), z["phnn_rk4/U_grad"]) < TOL


def test_canonical_euler_training_loss_gradients():
    """scripts/train_cartpole_phnn_canonical.py:83-180 with integrator='euler': position loss + velocity-reconstruction loss
    (q_dot_reconstructed = M^-1(M q_dot) evaluated along the predicted trajectory)"""
    from phnn_mpc_b200.training import trainable_rollout
    z, xb, ub, dt = _load()
    model = _model("canonical", z)
    y0 = xb[:, 0, :].clone().requires_grad_(True)
    traj = trainable_rollout(model, y0, ub[:, :-1, :], dt, "euler")
    assert rel_err(traj.detach().numpy(), z["canon_euler/traj"]) < TOL
    vel = []
    for t in range(xb.shape[1] - 1):
        qd = model.get_velocity_reconstruction(traj[:, t, :])
        vel.append(torch.sum((qd - xb[:, t, 2:]) ** 2, dim=1).mean())
    l_pos = torch.mean((traj[:, :, 0] - xb[:, :, 0]) ** 2) + torch.mean(1 - torch.cos(traj[:, :, 1] - xb[:, :, 1]))
    loss = l_pos + torch.mean(torch.stack(vel))
    assert abs(loss.item() - float(z["canon_euler/loss"])) < TOL * abs(float(z["canon_euler/loss"]))
    loss.backward()
    # the mass-matrix scalars are read with .item() by the reference on this path: zero gradient in the fixture, none here
    _check_param_grads(model, z, "canon_euler")


def test_rollout_vjp_is_linear_in_the_cotangent_and_matches_cost_grad():
    """size-independent properties on a larger batch: linearity in gtraj, and dL/dU for the MPC cost's own trajectory
    cotangent equals phnn_cost_grad's dJ/dU (same adjoint, two entry points)"""
    from phnn_mpc_b200 import ops
    from phnn_mpc_b200.packing import pack_of
    from phnn_mpc_b200.training import rollout_vjp
    z, xb, ub, dt = _load()
    model = _model("phnn", z)
    pk = pack_of(model)
    g = torch.Generator().manual_seed(5)
    B, T = 300, 12
    x0 = ((torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).cuda()
    U = ((torch.rand(B, T, 1, generator=g) * 2 - 1) * 5).cuda()
    g1 = torch.randn(B, T + 1, 4, generator=g).cuda()
    g2 = torch.randn(B, T + 1, 4, generator=g).cuda()
    want = {k: tuple(p.shape) for k, p in model.named_parameters()}
    a = rollout_vjp(pk, x0, U, g1, dt, "rk4", want)
    b = rollout_vjp(pk, x0, U, g2, dt, "rk4", want)
    c = rollout_vjp(pk, x0, U, 2.0 * g1 - 0.5 * g2, dt, "rk4", want)
    assert set(a[2]) == {k for k in want if k != "H_net.net.4.bias"}      # the output bias of H_net does not enter the dynamics
    for k in a[2]:
        comb = 2.0 * a[2][k] - 0.5 * b[2][k]
        assert (comb - c[2][k]).abs().max() <= 2e-4 * max(1e-6, comb.abs().max().item()), k
    assert (2.0 * a[1] - 0.5 * b[1] - c[1]).abs().max() <= 2e-4 * c[1].abs().max()
    # MPC cost sum_t e^T Q e (no control term, no clamp): gtraj = 2 Q (x_t - x*)
    Q = torch.diag(torch.tensor([10.0, 200.0, 1.0, 10.0]))
    traj, _ = ops.rollout(pk.handle, x0, U, dt, 1, 0)
    gt = 2.0 * traj @ Q.cuda()
    _, dU, _ = rollout_vjp(pk, x0, U, gt, dt, "rk4", {})
    _, gJ, _ = ops.cost_grad(pk.handle, x0, U, dt, 1, Q, torch.zeros(1, 1), torch.zeros(4), False, 0.0, 0.0, None, None, 1000.0, True, False)
    assert rel_err(dU.cpu().numpy(), gJ.cpu().numpy()) < 1e-5
