"""CPU: the C restatement (oracle/) against the reference's own outputs (tests/golden/).

Tolerances: float32 oracle vs float32 reference -- 2e-5 relative to the largest entry (both are
FP32 evaluations in different summation orders; SURVEY.md fact 10 puts the FP32 noise floor at
~1e-6); float64 oracle vs float32 reference -- same bound.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle.phnn_oracle import OracleModel

TOL = 2e-5
# canonical_constM: the canonical pHNN with MassMatrixNetwork(mass_type='constant') (src/mass_matrix.py:15-216), built by the
# reference's own constructor branch (src/pHNN_canonical.py:79-86)
# cartpole_h128_dropout: MLPs built with dropout = 0.1 (Linear layers at net.0 / net.3 / net.6), recorded in eval mode
KINDS = {"pendulum": "phnn", "cartpole_h128": "phnn", "cartpole_h256": "phnn", "canonical": "canonical",
         "canonical_constM": "canonical", "cartpole_h128_dropout": "phnn", "canonical_nobias": "canonical"}   # nobias: H_mlp bias: false


@pytest.mark.parametrize("name", list(KINDS))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_forward_and_vjp(name, dtype):
    z, sd = load_golden(name)
    M = OracleModel(sd, KINDS[name], dtype)
    dx, H = M.forward(z["rand_x"], z["rand_u"])
    assert rel_err(dx, z["rand_dx"]) < TOL
    assert rel_err(H, z["rand_H"]) < TOL
    xb, ub = M.vjp(z["rand_x"], z["rand_u"], z["rand_v"])
    assert rel_err(xb, z["rand_gx"]) < TOL
    assert rel_err(ub, z["rand_gu"]) < TOL


def test_pendulum_anchors():
    """SURVEY.md Appendix C known answers from the shipped pendulum weights."""
    z, sd = load_golden("pendulum")
    M = OracleModel(sd, "phnn")
    dx, H = M.forward(z["anchor_x"], z["anchor_u"])
    np.testing.assert_allclose(dx, [[0.181520462, -8.375887871], [0.221681133, 8.736040115]], rtol=2e-6)
    np.testing.assert_allclose(H, [-14.445344925, -10.949809074], rtol=2e-6)
    U10 = np.repeat(z["anchor_u"][:, None, :], 10, 1)
    for integ in ("rk4", "euler"):
        tr, en = M.rollout(z["anchor_x"], U10, 0.05, integ, energy_mode=1)
        assert rel_err(tr, z["anchor_traj_" + integ]) < TOL
        assert rel_err(en, z["anchor_en_" + integ]) < TOL
        tr2, en2 = M.rollout(z["anchor_x"], U10, 0.05, integ, energy_mode=2)
        assert rel_err(tr2, z["anchor_traj2_" + integ]) < TOL
        assert rel_err(en2, z["anchor_en2_" + integ]) < TOL
    tr = M.rollout(z["anchor_x"], U10, 0.05, "rk4")
    np.testing.assert_allclose(tr[:, -1], [[0.300435662, -2.837836504], [-0.848876774, 4.260723591]], rtol=1e-5)
    # J = sum traj^2 is a quadratic cost with Q = I, R = 0, target 0
    C = M.cost_struct(np.eye(2), np.zeros((1, 1)), np.zeros(2))
    J, g = M.cost_grad(C, z["anchor_x"], U10, 0.05, "rk4")
    assert abs(J.sum() - 148.06396484) < 2e-3
    assert rel_err(g, z["anchor_dJdU"]) < TOL


@pytest.mark.parametrize("integ", ["rk4", "euler"])
def test_pendulum_cfg2_rollout(integ):
    z, sd = load_golden("pendulum")
    M = OracleModel(sd, "phnn")
    tr, en = M.rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ, energy_mode=1)
    # 100 chained steps: horizon tolerance 1e-4 (north_star)
    assert rel_err(tr, z["cfg2_traj_" + integ]) < 1e-4
    assert rel_err(en, z["cfg2_en_" + integ]) < 1e-4


@pytest.mark.parametrize("name", list(KINDS))
@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_mpc_composition(name, integ):
    z, sd = load_golden(name)
    M = OracleModel(sd, KINDS[name])
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    C = M.cost_struct(z["mpc_Q"], z["mpc_R"], z["mpc_xt"], lo, hi)
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    J, g, tr = M.cost_grad(C, z["mpc_x0"], z["mpc_U0"], dt, integ, want_traj=True)
    p = "mpc_%s_" % integ
    assert rel_err(J, z[p + "hist"][0]) < TOL
    assert rel_err(g, z[p + "grad0"]) < 5e-5
    if p + "traj0" in z.files:
        assert rel_err(tr, z[p + "traj0"]) < TOL
    # clamped-out controls get exactly zero gradient
    out = (z["mpc_U0"] < lo) | (z["mpc_U0"] > hi)
    assert out.any() and np.all(g[out] == 0)
    iters = z[p + "hist"].shape[0]
    for mode, key in (("last", "U_last"), ("best", "U_best")):
        U, hist, best = M.mpc_solve(C, z["mpc_x0"], z["mpc_U0"], dt, integ, lr=lr, iters=iters, return_mode=mode)
        assert rel_err(hist, z[p + "hist"]) < 1e-4
        # controls: absolute tolerance scaled by lr (Adam steps are ~lr*sign(g) early on)
        assert np.abs(U - z[p + key]).max() < 0.02 * lr + 1e-5
        assert rel_err(best, z[p + "best"]) < 1e-4


def test_controller_cfg1():
    """MPCController.compute_control, B=1, YAML parameters (BASELINE config 1)."""
    z, sd = load_golden("cartpole_h128")
    M = OracleModel(sd, "phnn")
    C = M.cost_struct([10.0, 200.0, 1.0, 10.0], 0.01, np.zeros(4), -15.0, 15.0)
    x = z["ctrl_x"].astype(np.float32)
    U, _, _ = M.mpc_solve(C, x, np.zeros((3, 20, 1)), 0.02, "euler", lr=0.015, iters=30, return_mode="last")
    assert np.abs(U[:, 0] - z["ctrl_u"]).max() < 0.02 * 0.015
    Cb = M.cost_struct([10.0, 200.0, 1.0, 10.0], 0.01, np.zeros(4), -15.0, 15.0,
                       x_min=[-0.2, -0.05, -0.1, -0.2], x_max=[0.2, 0.05, 0.1, 0.2])
    J, = M.cost_grad(Cb, x[1:2], np.zeros((1, 8, 1)), 0.02, "euler", want_grad=False)
    assert abs(J[0] - z["ctrlb_cost0"]) / z["ctrlb_cost0"] < TOL
    Ub, _, _ = M.mpc_solve(Cb, x, np.zeros((3, 8, 1)), 0.02, "euler", lr=0.015, iters=6, return_mode="last")
    assert np.abs(Ub[:, 0] - z["ctrlb_u"]).max() < 0.02 * 0.015


def test_controller_cfg3():
    """MPCControllerCanonical.control, cold and warm start (BASELINE config 3, B=1)."""
    z, sd = load_golden("canonical")
    M = OracleModel(sd, "canonical")
    C = M.cost_struct([0.0, 1000.0, 0.0, 100.0], [1e-4], np.zeros(4), -30.0, 30.0)
    x = z["ctrl_x"].astype(np.float32)
    U1, hist1, _ = M.mpc_solve(C, x, np.zeros((3, 10, 1)), 0.02, "euler", lr=0.03, iters=50, return_mode="best")
    assert np.abs(U1 - z["ctrl_seq"][:, 0]).max() < 0.05 * 0.03
    assert rel_err(hist1.T, z["ctrl_costs"][:, 0]) < 1e-4
    warm = np.concatenate([z["ctrl_seq"][:, 0, 1:], np.zeros((3, 1, 1), np.float32)], 1)
    x2 = (z["ctrl_x"] + 0.01).astype(np.float32)
    U2, hist2, _ = M.mpc_solve(C, x2, warm, 0.02, "euler", lr=0.03, iters=50, return_mode="best")
    assert np.abs(U2 - z["ctrl_seq"][:, 1]).max() < 0.05 * 0.03
    assert rel_err(hist2.T, z["ctrl_costs"][:, 1]) < 1e-4


def test_oracle_cfg4_shape_vs_reference():
    """the benchmarked job's shape (cfg4 weights, bench.py's first 64 instances, H=50, RK4, 20 Adam iterations) recorded
    from the reference itself (make_golden.gen_cfg4_shape): cost history, dJ/dU at iteration 0, controls, best cost"""
    from oracle.phnn_oracle import OracleModel, set_threads
    import os
    z, _ = load_golden("cfg4_shape")
    _, sd = load_golden("cartpole_h256")
    set_threads(os.cpu_count() or 1)
    M = OracleModel(sd, "phnn")
    C = M.cost_struct(z["Q"], z["R"], z["xt"], float(z["bounds"][0]), float(z["bounds"][1]))
    B, H, iters = z["x0"].shape[0], int(z["H"]), int(z["iters"])
    U0 = np.zeros((B, H, 1), np.float32)
    U, hist, best = M.mpc_solve(C, z["x0"], U0, float(z["dt"]), "rk4", lr=float(z["lr"]), iters=iters)
    assert rel_err(hist, z["rk4_hist"]) < 1e-4 and rel_err(best, z["rk4_best"]) < 1e-4
    assert np.abs(U - z["rk4_U_last"]).max() < 0.02 * float(z["lr"]) + 1e-5
    _, g0 = M.cost_grad(C, z["x0"], U0, float(z["dt"]), "rk4")
    assert rel_err(g0, z["rk4_grad0"]) < 1e-4


@pytest.mark.parametrize("H", [100, 200])
def test_oracle_cfg5_shape_vs_reference(H):
    """BASELINE cfg5 horizons (cfg4 weights, bench.py's first 16 instances, H = 100 / 200, RK4, 4 Adam iterations) recorded
    from the reference itself (make_golden.gen_cfg5_shape).  Long horizons amplify FP32 rounding (the cart-pole model is
    unstable): stated bounds cost 1e-4 * H/50, dJ/dU 5e-4 * H/50, controls 0.05 * lr."""
    from oracle.phnn_oracle import OracleModel, set_threads
    import os
    z, _ = load_golden("cfg5_shape")
    _, sd = load_golden("cartpole_h256")
    set_threads(os.cpu_count() or 1)
    M = OracleModel(sd, "phnn")
    C = M.cost_struct(z["Q"], z["R"], z["xt"], float(z["bounds"][0]), float(z["bounds"][1]))
    B, iters = z["x0"].shape[0], int(z["iters"])
    U0 = np.zeros((B, H, 1), np.float32)
    p = "h%d_" % H
    tol = 1e-4 * H / 50.0
    U, hist, best = M.mpc_solve(C, z["x0"], U0, float(z["dt"]), "rk4", lr=float(z["lr"]), iters=iters)
    assert rel_err(hist, z[p + "hist"]) < tol and rel_err(best, z[p + "best"]) < tol
    assert np.abs(U - z[p + "U_last"]).max() < 0.05 * float(z["lr"])
    _, g0 = M.cost_grad(C, z["x0"], U0, float(z["dt"]), "rk4")
    assert rel_err(g0, z[p + "grad0"]) < 5 * tol
