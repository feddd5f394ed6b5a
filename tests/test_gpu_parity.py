"""GPU parity: the CUDA path (through the C ABI / torch.library ops) against
(a) the golden vectors recorded from the reference's PyTorch-autograd implementation and
(b) the CPU oracle on the same seeded inputs.

Tolerances (FP32 path, north_star: 1e-5 per step, 1e-4 over a horizon; all relative to the
largest entry of the compared tensor):
  single evaluation (dx, H, VJP) ......... 1e-5
  rollouts / cost / dJdU over a horizon .. 1e-4
  controls after Adam .................... abs 0.02*lr + 1e-5 (early Adam steps are ~lr*sign(g))
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

# canonical_constM: the canonical pHNN with MassMatrixNetwork(mass_type='constant') (src/mass_matrix.py:15-216), built by the
# reference's own constructor branch (src/pHNN_canonical.py:79-86)
# cartpole_h128_dropout: MLPs built with dropout = 0.1 (Linear layers at net.0 / net.3 / net.6), recorded in eval mode
KINDS = {"pendulum": "phnn", "cartpole_h128": "phnn", "cartpole_h256": "phnn", "canonical": "canonical",
         "canonical_constM": "canonical", "cartpole_h128_dropout": "phnn", "canonical_nobias": "canonical"}   # nobias: H_mlp bias: false
STEP_TOL = 1e-5
HORIZON_TOL = 1e-4


@pytest.fixture(scope="module")
def env():
    from phnn_mpc_b200 import ops
    from phnn_mpc_b200.packing import PackedModel
    assert torch.cuda.is_available()
    packs = {}

    def get(name):
        if name not in packs:
            z, sd = load_golden(name)
            pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, KINDS[name])
            if pk.get_option("latency_max_batch") > 0:
                pk.set_option("latency_max_batch", 0)      # these tests exercise the batched FP32-FMA kernel
            if pk.get_option("tensor_mode") > 0:
                pk.set_option("tensor_mode", 0)
            packs[name] = (z, sd, pk)
        return packs[name]

    return ops, get


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def cost_args(z):
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    return (torch.from_numpy(z["mpc_Q"]), torch.from_numpy(z["mpc_R"]), torch.from_numpy(z["mpc_xt"]), True, lo, hi,
            None, None, 1000.0)


@pytest.mark.parametrize("name", list(KINDS))
def test_forward_vjp_golden(env, name):
    ops, get = env
    z, sd, pk = get(name)
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
    assert rel_err(dx.cpu().numpy(), z["rand_dx"]) < STEP_TOL
    assert rel_err(H.cpu().numpy(), z["rand_H"]) < STEP_TOL
    xb, ub = ops.vjp(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]), cu(z["rand_v"]))
    assert rel_err(xb.cpu().numpy(), z["rand_gx"]) < STEP_TOL
    assert rel_err(ub.cpu().numpy(), z["rand_gu"]) < STEP_TOL


def test_forward_wide_states(env):
    """states over the range of data/cartpole_training_data.pt (|x| up to ~20)"""
    ops, get = env
    for name in ("cartpole_h128", "cartpole_h256"):
        z, sd, pk = get(name)
        dx, H = ops.forward(pk.handle, cu(z["wide_x"]), cu(z["rand_u"][:16]))
        assert rel_err(dx.cpu().numpy(), z["wide_dx"]) < STEP_TOL
        assert rel_err(H.cpu().numpy(), z["wide_H"]) < STEP_TOL
        xb, _ = ops.vjp(pk.handle, cu(z["wide_x"]), cu(z["rand_u"][:16]), cu(z["rand_v"][:16]))
        assert rel_err(xb.cpu().numpy(), z["wide_gx"]) < STEP_TOL


@pytest.mark.parametrize("B", [1, 31, 33, 100, 1000])
def test_forward_ragged_batches_vs_oracle(env, B):
    from oracle.phnn_oracle import OracleModel
    ops, get = env
    for name in KINDS:
        z, sd, pk = get(name)
        M = OracleModel(sd, KINDS[name])
        rng = np.random.default_rng(B)
        x = rng.normal(size=(B, M.n)).astype(np.float32)
        u = rng.normal(size=(B, 1)).astype(np.float32) * 3
        v = rng.normal(size=(B, M.n)).astype(np.float32)
        dx, H = ops.forward(pk.handle, cu(x), cu(u))
        dxo, Ho = M.forward(x, u)
        assert rel_err(dx.cpu().numpy(), dxo) < STEP_TOL
        assert rel_err(H.cpu().numpy(), Ho) < STEP_TOL
        xb, ub = ops.vjp(pk.handle, cu(x), cu(u), cu(v))
        xbo, ubo = M.vjp(x, u, v)
        assert rel_err(xb.cpu().numpy(), xbo) < STEP_TOL
        assert rel_err(ub.cpu().numpy(), ubo) < STEP_TOL


def test_empty_batch(env):
    ops, get = env
    z, sd, pk = get("cartpole_h128")
    dx, H = ops.forward(pk.handle, torch.empty(0, 4, device="cuda"), torch.empty(0, 1, device="cuda"))
    assert dx.shape == (0, 4) and H.shape == (0,)


def test_pendulum_anchor_rollouts(env):
    ops, get = env
    z, sd, pk = get("pendulum")
    U10 = np.repeat(z["anchor_u"][:, None, :], 10, 1)
    for integ, iid in (("rk4", 1), ("euler", 0)):
        tr, en = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, iid, 1)
        assert rel_err(tr.cpu().numpy(), z["anchor_traj_" + integ]) < HORIZON_TOL
        assert rel_err(en.cpu().numpy(), z["anchor_en_" + integ]) < HORIZON_TOL
        tr2, en2 = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, iid, 2)
        assert rel_err(tr2.cpu().numpy(), z["anchor_traj2_" + integ]) < HORIZON_TOL
        assert rel_err(en2.cpu().numpy(), z["anchor_en2_" + integ]) < HORIZON_TOL
    tr, _ = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, 1, 0)
    np.testing.assert_allclose(tr[:, -1].cpu().numpy(), [[0.300435662, -2.837836504], [-0.848876774, 4.260723591]],
                               rtol=2e-5)


@pytest.mark.parametrize("integ", ["rk4", "euler"])
def test_pendulum_cfg2_rollout(env, integ):
    """BASELINE config 2 inputs (first 64 instances), 100 chained steps, against the reference's
    own rollout.  This horizon amplifies rounding: two FP32 evaluations that differ only in
    summation order (reference vs the CPU FP32 oracle) are 7e-5 apart and the FP32 oracle is
    1e-4 from its FP64 twin, so the bound vs the reference is 2e-4 here, and the kernel must
    also be no further from FP64 than twice the CPU FP32 oracle is."""
    from oracle.phnn_oracle import OracleModel
    ops, get = env
    z, sd, pk = get("pendulum")
    tr, en = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, {"euler": 0, "rk4": 1}[integ], 1)
    tr, en = tr.cpu().numpy(), en.cpu().numpy()
    assert rel_err(tr, z["cfg2_traj_" + integ]) < 2e-4
    assert rel_err(en, z["cfg2_en_" + integ]) < 2e-4
    o64 = OracleModel(sd, "phnn", np.float64).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    o32 = OracleModel(sd, "phnn", np.float32).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    assert rel_err(tr, o64) < max(2 * rel_err(o32, o64), 5e-5)
    # the first 10 steps are within the per-step tolerance
    assert rel_err(tr[:, :11], z["cfg2_traj_" + integ][:, :11]) < STEP_TOL


@pytest.mark.parametrize("name", list(KINDS))
@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_cost_grad_and_solve_golden(env, name, integ):
    ops, get = env
    z, sd, pk = get(name)
    iid = {"euler": 0, "rk4": 1}[integ]
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    ca = cost_args(z)
    p = "mpc_%s_" % integ
    cost, g, tr = ops.cost_grad(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, True, True)
    assert rel_err(cost.cpu().numpy(), z[p + "hist"][0]) < HORIZON_TOL
    assert rel_err(g.cpu().numpy(), z[p + "grad0"]) < HORIZON_TOL
    if p + "traj0" in z.files:
        assert rel_err(tr.cpu().numpy(), z[p + "traj0"]) < HORIZON_TOL
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    out = (z["mpc_U0"] < lo) | (z["mpc_U0"] > hi)
    assert out.any() and np.all(g.cpu().numpy()[out] == 0)
    iters = z[p + "hist"].shape[0]
    for mode, key in ((0, "U_last"), (1, "U_best")):
        U, hist, best = ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, lr, 0.9, 0.999, 1e-8,
                                      iters, mode, True)
        assert rel_err(hist.cpu().numpy(), z[p + "hist"]) < HORIZON_TOL
        assert np.abs(U.cpu().numpy() - z[p + key]).max() < 0.02 * lr + 1e-5
        assert rel_err(best.cpu().numpy(), z[p + "best"]) < HORIZON_TOL


@pytest.mark.parametrize("name,H,integ", [("cartpole_h256", 50, "rk4"), ("cartpole_h256", 50, "euler"),
                                          ("cartpole_h128", 20, "euler"), ("canonical", 10, "euler"),
                                          ("pendulum", 30, "rk4")])
def test_solve_vs_oracle_batched(env, name, H, integ):
    """BASELINE config 3/4-shaped work on a 96-instance sub-sample against the CPU oracle."""
    from oracle.phnn_oracle import OracleModel
    ops, get = env
    z, sd, pk = get(name)
    M = OracleModel(sd, KINDS[name])
    B, iters = 96, 4
    g = torch.Generator().manual_seed(7)
    n = M.n
    scale = torch.tensor([1.0, 0.3, 0.5, 0.5][:n])
    x0 = ((torch.rand(B, n, generator=g) * 2 - 1) * scale).numpy()
    U0 = ((torch.rand(B, H, 1, generator=g) * 2 - 1) * 3).numpy()
    Q = np.diag([10.0, 200.0, 1.0, 10.0][:n]).astype(np.float32)
    R = np.array([[0.01]], np.float32)
    xt = np.zeros(n, np.float32)
    C = M.cost_struct(Q, R, xt, -15.0, 15.0)
    Jo, go = M.cost_grad(C, x0, U0, 0.02, integ)
    iid = {"euler": 0, "rk4": 1}[integ]
    ca = (torch.from_numpy(Q), torch.from_numpy(R), torch.from_numpy(xt), True, -15.0, 15.0, None, None, 1000.0)
    cost, gg, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), 0.02, iid, *ca, True, False)
    assert rel_err(cost.cpu().numpy(), Jo) < HORIZON_TOL
    assert rel_err(gg.cpu().numpy(), go) < HORIZON_TOL
    Uo, histo, besto = M.mpc_solve(C, x0, U0, 0.02, integ, lr=0.015, iters=iters, return_mode="last")
    U, hist, best = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, iid, *ca, 0.015, 0.9, 0.999, 1e-8, iters, 0, True)
    assert rel_err(hist.cpu().numpy(), histo) < HORIZON_TOL
    assert np.abs(U.cpu().numpy() - Uo).max() < 0.02 * 0.015 + 1e-5


def test_state_barrier_cost(env):
    """optional soft state bounds of MPCController (src/mpc_controller.py:96-107)"""
    ops, get = env
    z, sd, pk = get("cartpole_h128")
    x = cu(z["ctrl_x"][1:2])
    Q = torch.diag(torch.tensor([10.0, 200.0, 1.0, 10.0]))
    R = torch.tensor([[0.01]])
    xmin = torch.tensor([-0.2, -0.05, -0.1, -0.2])
    xmax = torch.tensor([0.2, 0.05, 0.1, 0.2])
    cost, _, _ = ops.cost_grad(pk.handle, x, torch.zeros(1, 8, 1, device="cuda"), 0.02, 0, Q, R, torch.zeros(4), True,
                               -15.0, 15.0, xmin, xmax, 1000.0, False, False)
    assert abs(cost.item() - float(z["ctrlb_cost0"])) / float(z["ctrlb_cost0"]) < HORIZON_TOL
    U, _, _ = ops.mpc_solve(pk.handle, cu(z["ctrl_x"]), torch.zeros(3, 8, 1, device="cuda"), 0.02, 0, Q, R,
                            torch.zeros(4), True, -15.0, 15.0, xmin, xmax, 1000.0, 0.015, 0.9, 0.999, 1e-8, 6, 0, False)
    assert np.abs(U[:, 0].cpu().numpy() - z["ctrlb_u"]).max() < 0.02 * 0.015


def test_unknown_integrator_raises(env):
    ops, get = env
    z, sd, pk = get("cartpole_h128")
    with pytest.raises(ValueError):
        ops.rollout(pk.handle, cu(z["rand_x"]), torch.zeros(32, 3, 1, device="cuda"), 0.02, 7, 0)


# ---------------------------------------------------------------------------------------------
# tcgen05 / TMEM kernel (cart-pole pHNN, hidden 128 and 256).  tensor_mode 3 = 3xTF32 error
# compensation and tensor_mode 2 (default) = TF32 + one BF16 correction product; both must meet the SAME FP32
# tolerances as the FP32-FMA kernel; tensor_mode 1 =
# plain TF32, stated looser tolerance (TF32 operand rounding, SURVEY.md fact 10: ~4e-5 rollout,
# ~4e-4 dJ/dU): 2e-4 cost/rollout, 1e-3 gradient, controls 0.05*lr.
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tc_env(env):
    from phnn_mpc_b200.packing import PackedModel
    ops, get = env
    packs = {}

    def get_tc(name, mode):
        key = (name, mode)
        if key not in packs:
            z, sd = load_golden(name)
            pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, KINDS[name])
            pk.set_option("tensor_min_batch", 0)
            pk.set_option("latency_max_batch", 0)
            pk.set_option("tensor_mode", mode)
            assert pk.get_option("tensor_mode") == mode
            packs[key] = (z, sd, pk)
        return packs[key]

    return ops, get_tc


@pytest.mark.parametrize("name", ["cartpole_h128", "cartpole_h256", "canonical", "canonical_constM", "cartpole_h128_dropout", "canonical_nobias"])
@pytest.mark.parametrize("mode", [4, 3, 2, 1, 5])
def test_tc_forward_rollout_costgrad_solve_golden(tc_env, name, mode):
    # modes 4 (3 x FP16 hi/lo), 3 (3xTF32) and 2 (TF32 + BF16 correction product) are held to the FP32 tolerances; mode 1
    # (plain TF32) and mode 5 (one FP16 product: 11-bit operands like TF32) to the stated looser ones
    ops, get_tc = tc_env
    z, sd, pk = get_tc(name, mode)
    step_tol, hor_tol, grad_tol, u_fac = (STEP_TOL, HORIZON_TOL, HORIZON_TOL, 0.02) if mode not in (1, 5) else (2e-3, 2e-4, 1e-3, 0.05)
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
    assert rel_err(dx.cpu().numpy(), z["rand_dx"]) < step_tol
    assert rel_err(H.cpu().numpy(), z["rand_H"]) < step_tol
    ca = cost_args(z)
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    for integ, iid in (("euler", 0), ("rk4", 1)):
        p = "mpc_%s_" % integ
        tr, _ = ops.rollout(pk.handle, cu(z["mpc_x0"]), cu(np.clip(z["mpc_U0"], lo, hi)), dt, iid, 0)
        cost, g, tr2 = ops.cost_grad(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, True, True)
        assert rel_err(cost.cpu().numpy(), z[p + "hist"][0]) < hor_tol
        assert rel_err(g.cpu().numpy(), z[p + "grad0"]) < grad_tol
        if p + "traj0" in z.files:
            assert rel_err(tr.cpu().numpy(), z[p + "traj0"]) < hor_tol
            assert rel_err(tr2.cpu().numpy(), z[p + "traj0"]) < hor_tol
        else:
            assert rel_err(tr.cpu().numpy(), tr2.cpu().numpy()) < hor_tol
        out = (z["mpc_U0"] < lo) | (z["mpc_U0"] > hi)
        assert np.all(g.cpu().numpy()[out] == 0)
        iters = z[p + "hist"].shape[0]
        for rmode, key in ((0, "U_last"), (1, "U_best")):
            U, hist, best = ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, lr, 0.9, 0.999, 1e-8,
                                          iters, rmode, True)
            assert rel_err(hist.cpu().numpy(), z[p + "hist"]) < hor_tol
            assert np.abs(U.cpu().numpy() - z[p + key]).max() < u_fac * lr + 1e-5
            assert rel_err(best.cpu().numpy(), z[p + "best"]) < hor_tol


@pytest.mark.parametrize("mode", [4, 3])
@pytest.mark.parametrize("B", [1, 127, 129, 700])
@pytest.mark.parametrize("name", ["cartpole_h256", "canonical"])
def test_tc_ragged_tiles_vs_oracle(tc_env, B, name, mode):
    """partially filled 128-instance tiles, energies in both orderings, against the CPU oracle"""
    from oracle.phnn_oracle import OracleModel
    ops, get_tc = tc_env
    z, sd, pk = get_tc(name, mode)
    M = OracleModel(sd, KINDS[name])
    rng = np.random.default_rng(B)
    x = (rng.uniform(-1, 1, size=(B, 4)) * [1.0, 0.3, 0.5, 0.5]).astype(np.float32)
    U = rng.uniform(-5, 5, size=(B, 6, 1)).astype(np.float32)
    dx, H = ops.forward(pk.handle, cu(x), cu(U[:, 0]))
    dxo, Ho = M.forward(x, U[:, 0])
    # H is a cancelling sum of O(1) terms: normalise by at least 1 (a single tiny |H| would inflate rel_err)
    assert rel_err(dx.cpu().numpy(), dxo) < STEP_TOL
    assert np.abs(H.cpu().numpy() - Ho).max() < STEP_TOL * max(1.0, np.abs(Ho).max())
    for emode in (1, 2):
        tr, en = ops.rollout(pk.handle, cu(x), cu(U), 0.02, 1, emode)
        tro, eno = M.rollout(x, U, 0.02, "rk4", energy_mode=emode)
        assert rel_err(tr.cpu().numpy(), tro) < HORIZON_TOL and rel_err(en.cpu().numpy(), eno) < HORIZON_TOL


def test_tc_matches_fp32_kernel_and_oracle_cfg4_shape(tc_env, env):
    """BASELINE cfg4-shaped work (h=256, H=50, RK4) on 256 instances: tcgen05 kernel vs the CPU oracle and
    vs the FP32-FMA kernel."""
    from oracle.phnn_oracle import OracleModel
    ops, get_tc = tc_env
    _, get = env
    z, sd, pk_tc = get_tc("cartpole_h256", 3)
    _, _, pk_fp = get("cartpole_h256")
    M = OracleModel(sd, "phnn")
    B, H, iters = 256, 50, 3
    g = torch.Generator().manual_seed(11)
    x0 = ((torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).numpy()
    Q = np.diag([10.0, 200.0, 1.0, 10.0]).astype(np.float32)
    R = np.array([[0.01]], np.float32)
    ca = (torch.from_numpy(Q), torch.from_numpy(R), torch.zeros(4), True, -15.0, 15.0, None, None, 1000.0)
    U0 = np.zeros((B, H, 1), np.float32)
    C = M.cost_struct(Q, R, np.zeros(4), -15.0, 15.0)
    Uo, histo, _ = M.mpc_solve(C, x0, U0, 0.02, "rk4", lr=0.015, iters=iters)
    outs = []
    for pk in (pk_tc, pk_fp):
        U, hist, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, 0.015, 0.9, 0.999, 1e-8, iters, 0, True)
        assert rel_err(hist.cpu().numpy(), histo) < HORIZON_TOL
        assert np.abs(U.cpu().numpy() - Uo).max() < 0.02 * 0.015 + 1e-5
        outs.append(U.cpu().numpy())
    assert np.abs(outs[0] - outs[1]).max() < 0.02 * 0.015 + 1e-5


# ---------------------------------------------------------------------------------------------
# latency kernel (one CTA per instance; the default route for B <= 2 x SM count, hidden <= 128)
# ---------------------------------------------------------------------------------------------
LAT_MODELS = ["pendulum", "cartpole_h128", "canonical", "canonical_constM", "cartpole_h128_dropout", "canonical_nobias"]


@pytest.fixture(scope="module")
def lat_env(env):
    from phnn_mpc_b200.packing import PackedModel
    ops, _ = env
    packs = {}

    def get_lat(name):
        if name not in packs:
            z, sd = load_golden(name)
            pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, KINDS[name])
            assert pk.get_option("latency_max_batch") >= 32
            packs[name] = (z, sd, pk)
        return packs[name]

    return ops, get_lat


@pytest.mark.parametrize("name", LAT_MODELS)
def test_lat_forward_vjp_golden(lat_env, name):
    ops, get_lat = lat_env
    z, sd, pk = get_lat(name)
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
    assert rel_err(dx.cpu().numpy(), z["rand_dx"]) < STEP_TOL
    assert rel_err(H.cpu().numpy(), z["rand_H"]) < STEP_TOL
    xb, ub = ops.vjp(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]), cu(z["rand_v"]))
    assert rel_err(xb.cpu().numpy(), z["rand_gx"]) < STEP_TOL
    assert rel_err(ub.cpu().numpy(), z["rand_gu"]) < STEP_TOL
    # a single instance (the reference's own call shape)
    dx1, H1 = ops.forward(pk.handle, cu(z["rand_x"][:1]), cu(z["rand_u"][:1]))
    assert rel_err(dx1.cpu().numpy(), z["rand_dx"][:1]) < STEP_TOL


def test_lat_pendulum_rollouts(lat_env):
    ops, get_lat = lat_env
    z, sd, pk = get_lat("pendulum")
    U10 = np.repeat(z["anchor_u"][:, None, :], 10, 1)
    for integ, iid in (("rk4", 1), ("euler", 0)):
        for emode, key in ((1, "anchor_%s_" + integ), (2, "anchor_%s2_" + integ)):
            tr, en = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, iid, emode)
            assert rel_err(tr.cpu().numpy(), z[key % "traj"]) < HORIZON_TOL
            assert rel_err(en.cpu().numpy(), z[key % "en"]) < HORIZON_TOL
    tr, _ = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, 1, 0)
    assert rel_err(tr.cpu().numpy(), z["cfg2_traj_rk4"]) < 2e-4


@pytest.mark.parametrize("name", LAT_MODELS)
@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_lat_cost_grad_and_solve_golden(lat_env, name, integ):
    ops, get_lat = lat_env
    z, sd, pk = get_lat(name)
    iid = {"euler": 0, "rk4": 1}[integ]
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    ca = cost_args(z)
    p = "mpc_%s_" % integ
    cost, g, tr = ops.cost_grad(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, True, True)
    assert rel_err(cost.cpu().numpy(), z[p + "hist"][0]) < HORIZON_TOL
    assert rel_err(g.cpu().numpy(), z[p + "grad0"]) < HORIZON_TOL
    lo, hi = [float(v) for v in z["mpc_bounds"]]
    out = (z["mpc_U0"] < lo) | (z["mpc_U0"] > hi)
    assert np.all(g.cpu().numpy()[out] == 0)
    iters = z[p + "hist"].shape[0]
    for mode, key in ((0, "U_last"), (1, "U_best")):
        U, hist, best = ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, iid, *ca, lr, 0.9, 0.999, 1e-8,
                                      iters, mode, True)
        assert rel_err(hist.cpu().numpy(), z[p + "hist"]) < HORIZON_TOL
        assert np.abs(U.cpu().numpy() - z[p + key]).max() < 0.02 * lr + 1e-5
        assert rel_err(best.cpu().numpy(), z[p + "best"]) < HORIZON_TOL


def test_lat_matches_batched_kernel(lat_env, env):
    """same inputs through the latency kernel and the batched FP32 kernel"""
    ops, get_lat = lat_env
    _, get = env
    for name in LAT_MODELS:
        z, sd, pk_lat = get_lat(name)
        _, _, pk_fp = get(name)
        ca = cost_args(z)
        dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
        outs = [ops.mpc_solve(pk.handle, cu(z["mpc_x0"]), cu(z["mpc_U0"]), dt, 1, *ca, lr, 0.9, 0.999, 1e-8, 4, 0, True)
                for pk in (pk_lat, pk_fp)]
        assert rel_err(outs[0][1].cpu().numpy(), outs[1][1].cpu().numpy()) < HORIZON_TOL
        assert np.abs(outs[0][0].cpu().numpy() - outs[1][0].cpu().numpy()).max() < 0.02 * lr + 1e-5


@pytest.mark.parametrize("name,B", [("cartpole_h128", 300), ("canonical", 450), ("pendulum", 1500)])
def test_lat_stacked_instances_vs_oracle(lat_env, name, B):
    """batches between one and several instances per SM: the latency kernel stacks 2..8 instances on a CTA (shared
    weights, per-instance barriers, ragged last CTA); cost, dJ/dU and three Adam steps against the CPU oracle"""
    from oracle.phnn_oracle import OracleModel
    ops, get_lat = lat_env
    z, sd, pk = get_lat(name)
    assert B <= pk.get_option("latency_max_batch")
    M = OracleModel(sd, KINDS[name])
    n = z["mpc_x0"].shape[1]
    H = 12
    g = torch.Generator().manual_seed(B)
    scale = torch.tensor([1.0, 0.3, 0.5, 0.5][:n]) if n == 4 else torch.tensor([1.5, 1.0])
    x0 = ((torch.rand(B, n, generator=g) * 2 - 1) * scale).numpy()
    U0 = ((torch.rand(B, H, 1, generator=g) * 2 - 1) * 2).numpy()
    ca = cost_args(z)
    Q, R, xt = [t.numpy() for t in ca[:3]]
    lo, hi = ca[4], ca[5]
    C = M.cost_struct(Q, R, xt, lo, hi)
    dt, lr = float(z["mpc_dt"]), float(z["mpc_lr"])
    Jo, go = M.cost_grad(C, x0, U0, dt, "rk4")
    cost, gg, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), dt, 1, *ca, True, False)
    assert rel_err(cost.cpu().numpy(), Jo) < HORIZON_TOL
    assert rel_err(gg.cpu().numpy(), go) < HORIZON_TOL
    Uo, histo, _ = M.mpc_solve(C, x0, U0, dt, "rk4", lr=lr, iters=3)
    U, hist, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), dt, 1, *ca, lr, 0.9, 0.999, 1e-8, 3, 0, True)
    assert rel_err(hist.cpu().numpy(), histo) < HORIZON_TOL
    assert np.abs(U.cpu().numpy() - Uo).max() < 0.02 * lr + 1e-5


# ---------------------------------------------------------------------------------------------
# BASELINE cfg5 horizons and full-size properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [4, 2, 3])
@pytest.mark.parametrize("H", [100, 200])
def test_tc_long_horizons_vs_oracle(tc_env, H, mode):
    """cfg5 horizons (RK4, h=256) on a 128-instance sub-sample against the CPU oracle: cost, dJ/dU, two Adam steps;
    both FP32-level tensor schemes (2: TF32 + BF16 correction product, the default; 3: 3xTF32)."""
    from oracle.phnn_oracle import OracleModel
    ops, get_tc = tc_env
    z, sd, pk = get_tc("cartpole_h256", mode)
    M = OracleModel(sd, "phnn")
    B = 128
    g = torch.Generator().manual_seed(H)
    x0 = ((torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).numpy()
    U0 = ((torch.rand(B, H, 1, generator=g) * 2 - 1) * 2).numpy()
    Q = np.diag([10.0, 200.0, 1.0, 10.0]).astype(np.float32)
    R = np.array([[0.01]], np.float32)
    C = M.cost_struct(Q, R, np.zeros(4), -15.0, 15.0)
    ca = (torch.from_numpy(Q), torch.from_numpy(R), torch.zeros(4), True, -15.0, 15.0, None, None, 1000.0)
    Jo, go = M.cost_grad(C, x0, U0, 0.02, "rk4")
    cost, gg, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, True, False)
    # long horizons amplify rounding (the cart-pole model is unstable): 4*H steps of FP32 arithmetic.  Stated bounds
    # (DESIGN.md section 2): cost 1e-4 * H/50, dJ/dU 5e-4 * H/50 against the FP32 oracle, AND the FP64 tie-breaker: the
    # kernel may be no further from the FP64 oracle than twice the FP32 oracle itself is (floor: the H=50 tolerance)
    tol = HORIZON_TOL * (H / 50.0)
    assert rel_err(cost.cpu().numpy(), Jo) < tol
    assert rel_err(gg.cpu().numpy(), go) < 5 * tol
    M64 = OracleModel(sd, "phnn", np.float64)
    C64 = M64.cost_struct(Q, R, np.zeros(4), -15.0, 15.0)
    J64, g64 = M64.cost_grad(C64, x0, U0, 0.02, "rk4")
    assert rel_err(cost.cpu().numpy(), J64) < max(2 * rel_err(Jo, J64), HORIZON_TOL)
    assert rel_err(gg.cpu().numpy(), g64) < max(2 * rel_err(go, g64), HORIZON_TOL)
    Uo, histo, _ = M.mpc_solve(C, x0, U0, 0.02, "rk4", lr=0.015, iters=2)
    U, hist, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, 0.015, 0.9, 0.999, 1e-8, 2, 0, True)
    assert rel_err(hist.cpu().numpy(), histo) < tol
    assert np.abs(U.cpu().numpy() - Uo).max() < 0.05 * 0.015


@pytest.mark.parametrize("mode", [4, 2])
def test_full_size_properties_cfg4(tc_env, mode):
    """BASELINE cfg4 batch size (65 536 instances, h=256) through size-independent properties: sharding and
    permutation invariance (bit for bit: instances are independent and every reduction order is fixed), bounded
    controls, monotone best cost, zero-iteration solve."""
    ops, get_tc = tc_env
    z, sd, pk = get_tc("cartpole_h256", mode)
    B, H, iters = 65536, 8, 3
    g = torch.Generator().manual_seed(4)
    x0 = ((torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).cuda()
    U0 = ((torch.rand(B, H, 1, generator=g) * 2 - 1) * 20).cuda()
    Q = torch.diag(torch.tensor([10.0, 200.0, 1.0, 10.0]))
    ca = (Q, torch.tensor([[0.01]]), torch.zeros(4), True, -15.0, 15.0, None, None, 1000.0)
    solve = lambda x, U, it=iters, mode=1: ops.mpc_solve(pk.handle, x, U, 0.02, 1, *ca, 0.015, 0.9, 0.999, 1e-8, it, mode, True)
    U, hist, best = solve(x0, U0)
    assert torch.isfinite(U).all() and U.abs().max() <= 15.0
    assert torch.equal(best, hist.min(0).values)                       # best = min over the cost history
    # shards solved separately == the full batch (what the multi-GPU path relies on)
    cut = 40000                                                        # not a multiple of the 128-instance tile
    Ua, ha, ba = solve(x0[:cut].contiguous(), U0[:cut].contiguous())
    Ub, hb, bb = solve(x0[cut:].contiguous(), U0[cut:].contiguous())
    assert torch.equal(torch.cat([Ua, Ub]), U) and torch.equal(torch.cat([ba, bb]), best)
    assert torch.equal(torch.cat([ha, hb], 1), hist)
    # permutation of the instances permutes the results
    perm = torch.randperm(B, generator=g).cuda()
    Up, hp, bp = solve(x0[perm].contiguous(), U0[perm].contiguous())
    assert torch.equal(Up, U[perm]) and torch.equal(bp, best[perm])
    # zero iterations: the clamped initial guess (src/mpc_controller.py:203-207)
    Uz, _, _ = ops.mpc_solve(pk.handle, x0, U0, 0.02, 1, *ca, 0.015, 0.9, 0.999, 1e-8, 0, 0, False)
    assert torch.equal(Uz, U0.clamp(-15.0, 15.0))
    # last-iterate mode equals the Adam trajectory's end, and differs from best-iterate only where the last is not best
    Ul, hl, _ = solve(x0, U0, iters, 0)
    assert torch.equal(hl, hist)


def test_rollout_linearity_in_time_cfg2_size():
    """BASELINE cfg2 size (4096 pendulum instances, H=100, RK4): a rollout of T steps equals two chained rollouts of T/2
    (the kernel carries no hidden state between steps), and the energy orderings are consistent."""
    from phnn_mpc_b200 import ops
    from phnn_mpc_b200.packing import PackedModel
    z, sd = load_golden("pendulum")
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
    g = torch.Generator().manual_seed(1)
    B, T = 4096, 100
    x0 = torch.stack([(torch.rand(B, generator=g) * 2 - 1) * np.pi, torch.rand(B, generator=g) * 2 - 1], 1).cuda()
    U = (torch.rand(B, T, 1, generator=g) * 4 - 2).cuda()
    tr, en1 = ops.rollout(pk.handle, x0, U, 0.05, 1, 1)
    _, en2 = ops.rollout(pk.handle, x0, U, 0.05, 1, 2)
    tra, _ = ops.rollout(pk.handle, x0, U[:, :50].contiguous(), 0.05, 1, 0)
    trb, _ = ops.rollout(pk.handle, tra[:, -1].contiguous(), U[:, 50:].contiguous(), 0.05, 1, 0)
    assert torch.equal(tr[:, :51], tra) and torch.equal(tr[:, 50:], trb)
    assert torch.equal(en1[:, 1:], en2[:, :-1]) and torch.equal(en1[:, 0], en2[:, 0])
    assert torch.isfinite(tr).all()


def test_workspace_contract(tc_env):
    """phnn_workspace_bytes follows the documented layout (per-instance part + scheduler words + one tape region per
    SM for the tcgen05 route), shrinks to the per-instance part when the tcgen05 route is off, and a workspace that is
    too small is refused with PHNN_E_WORKSPACE instead of being overrun."""
    import ctypes
    from phnn_mpc_b200 import _lib
    ops, get_tc = tc_env
    z, sd, pk = get_tc("cartpole_h256", 2)
    L = _lib.lib()
    L.phnn_workspace_bytes.restype = ctypes.c_size_t
    L.phnn_workspace_bytes.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B, T, h, n = 1000, 7, 256, 4
    for integ, S in ((0, 1), (1, 4)):
        tiles = (B + 127) // 128
        persist = tiles * 128 * (3 * T + 1) * 4                      # Adam m, v, best controls, best cost per instance
        sched = -(-(persist + 4 * (tiles + 4)) // 128) * 128          # + work-stealing words
        tape = min(tiles, sms) * T * S * 3 * h * 128 * 4              # per-CTA activation tape
        scratch = min(tiles, sms) * T * S * (n + 16) * 128 * 4        # per-CTA stage states + R_net sums / grad H
        assert L.phnn_workspace_bytes(ctypes.c_void_p(pk.handle), B, T, integ) == sched + tape + scratch
    # 1 M instances x H = 200 (BASELINE cfg5) fit one 180 GB GPU with room: < 60 GB
    assert L.phnn_workspace_bytes(ctypes.c_void_p(pk.handle), 1 << 20, 200, 1) < 60 * (1 << 30)
    pk.set_option("tensor_mode", 0)
    try:
        assert L.phnn_workspace_bytes(ctypes.c_void_p(pk.handle), B, T, 1) == -(-(1024 * (T * 4 * n + 3 * T + 1) * 4) // 128) * 128
    finally:
        pk.set_option("tensor_mode", 2)
    x0 = torch.zeros(B, n, device="cuda")
    U = torch.zeros(B, T, 1, device="cuda")
    ca = cost_args(z)
    cd = ops._Cost(*ca)
    small = torch.empty(1024, device="cuda")
    cost = torch.empty(B, device="cuda")
    g = torch.empty_like(U)
    rc = L.phnn_cost_grad(ctypes.c_void_p(pk.handle), ctypes.byref(cd.desc), ctypes.c_void_p(x0.data_ptr()),
                          ctypes.c_void_p(U.data_ptr()), ctypes.c_void_p(cost.data_ptr()), ctypes.c_void_p(g.data_ptr()), None,
                          B, T, ctypes.c_double(0.02), 1, ctypes.c_void_p(small.data_ptr()), small.numel() * 4, None)
    assert rc == _lib.E_WORKSPACE
    assert b"workspace too small" in L.phnn_last_error()


@pytest.mark.parametrize("mode", [4, 2, 3])
def test_tc_wide_states_and_saturated_units(tc_env, mode):
    """the tensor-core schemes at the edge of their operand range: states over the range of the reference's training
    data (|x| up to ~20: most tanh units saturate, so a1 is +-1 and delta2 / s1 are tiny) - forward against the golden
    values recorded from the reference, cost and dJ/dU of a short horizon from those states against the CPU oracle"""
    from oracle.phnn_oracle import OracleModel
    ops, get_tc = tc_env
    for name in ("cartpole_h128", "cartpole_h256"):
        z, sd, pk = get_tc(name, mode)
        dx, H = ops.forward(pk.handle, cu(z["wide_x"]), cu(z["rand_u"][:16]))
        assert rel_err(dx.cpu().numpy(), z["wide_dx"]) < STEP_TOL
        assert rel_err(H.cpu().numpy(), z["wide_H"]) < STEP_TOL
        M = OracleModel(sd, "phnn")
        x0 = np.ascontiguousarray(z["wide_x"], np.float32)
        B, T = x0.shape[0], 6
        U0 = np.random.default_rng(5).uniform(-3, 3, size=(B, T, 1)).astype(np.float32)
        ca = cost_args(z)
        C = M.cost_struct(ca[0].numpy(), ca[1].numpy(), ca[2].numpy(), ca[4], ca[5])
        Jo, go = M.cost_grad(C, x0, U0, 0.02, "rk4")
        cost, gg, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, True, False)
        assert rel_err(cost.cpu().numpy(), Jo) < HORIZON_TOL
        assert rel_err(gg.cpu().numpy(), go) < HORIZON_TOL


# ---------------------------------------------------------------------------------------------
# the benchmarked configuration itself: tcgen05 tensor_mode 2, h=256, H=50, RK4, 20 Adam iterations
# ---------------------------------------------------------------------------------------------
def _bench_inputs(B):
    import bench
    return bench.make_inputs(65536, "phnn", 7)[:B].contiguous()


def _cfg4_cost():
    Q = np.diag([10.0, 200.0, 1.0, 10.0]).astype(np.float32)
    R = np.array([[0.01]], np.float32)
    ca = (torch.from_numpy(Q), torch.from_numpy(R), torch.zeros(4), True, -15.0, 15.0, None, None, 1000.0)
    return Q, R, ca


@pytest.mark.parametrize("mode", [4, 2])
def test_benchmarked_config_vs_oracle(tc_env, mode):
    """bench.py's default job on 256 of its own instances (the first 256 of rank 0): tensor_mode 2, hidden 256, H=50,
    RK4, 20 Adam iterations, cold start -- cost history, dJ/dU at iteration 0 and at iteration 19, final controls,
    against the CPU oracle (src/mpc_controller.py:143-209 composed with src/integrators.py:192-258)."""
    from oracle.phnn_oracle import OracleModel, set_threads
    import os
    ops, get_tc = tc_env
    z, sd, pk = get_tc("cartpole_h256", mode)
    set_threads(os.cpu_count() or 1)
    M = OracleModel(sd, "phnn")
    B, H, iters, lr = 256, 50, 20, 0.015
    x0 = _bench_inputs(B).numpy()
    U0 = np.zeros((B, H, 1), np.float32)
    Q, R, ca = _cfg4_cost()
    C = M.cost_struct(Q, R, np.zeros(4), -15.0, 15.0)
    Uo, histo, besto = M.mpc_solve(C, x0, U0, 0.02, "rk4", lr=lr, iters=iters)
    pk.set_option("tensor_pair", 1)
    try:
        U, hist, best = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, lr, 0.9, 0.999, 1e-8, iters, 0, True)
        torch.cuda.synchronize()
    finally:
        pk.set_option("tensor_pair", 0)
    assert rel_err(hist.cpu().numpy(), histo) < HORIZON_TOL
    assert rel_err(best.cpu().numpy(), besto) < HORIZON_TOL
    assert np.abs(U.cpu().numpy() - Uo).max() < 0.02 * lr + 1e-5
    # dJ/dU at iteration 0 (U = 0) and at iteration 19 (the iterate the 20th gradient is taken at)
    J0, g0 = M.cost_grad(C, x0, U0, 0.02, "rk4")
    c0, gg0, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, True, False)
    assert rel_err(c0.cpu().numpy(), J0) < HORIZON_TOL and rel_err(gg0.cpu().numpy(), g0) < HORIZON_TOL
    U19, _, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, lr, 0.9, 0.999, 1e-8, iters - 1, 0, False)
    assert U19.abs().max() < 15.0                                  # no clamping: U19 is the raw Adam iterate
    J19, g19 = M.cost_grad(C, x0, U19.cpu().numpy(), 0.02, "rk4")
    c19, gg19, _ = ops.cost_grad(pk.handle, cu(x0), U19, 0.02, 1, *ca, True, False)
    assert rel_err(c19.cpu().numpy(), J19) < HORIZON_TOL and rel_err(gg19.cpu().numpy(), g19) < HORIZON_TOL
    assert rel_err(c19.cpu().numpy(), histo[iters - 1]) < HORIZON_TOL   # and it is the 20th entry of the history


@pytest.mark.parametrize("mode", [4, 2, 3])
def test_cfg4_shape_golden_reference(tc_env, mode):
    """the same job on 64 instances against the fixture recorded from the REFERENCE ITSELF (one hop, no oracle):
    tests/golden/cfg4_shape.npz = make_golden.gen_cfg4_shape (reference PyTorch autograd + torch.optim.Adam)."""
    ops, get_tc = tc_env
    _, sd, pk = get_tc("cartpole_h256", mode)
    z, _ = load_golden("cfg4_shape")
    B, H, iters, lr = 64, int(z["H"]), int(z["iters"]), float(z["lr"])
    x0 = _bench_inputs(B).numpy()
    assert np.array_equal(x0, z["x0"])                            # the fixture holds bench.py's own first 64 instances
    U0 = np.zeros((B, H, 1), np.float32)
    _, _, ca = _cfg4_cost()
    U, hist, best = ops.mpc_solve(pk.handle, cu(x0), cu(U0), float(z["dt"]), 1, *ca, lr, 0.9, 0.999, 1e-8, iters, 0, True)
    assert rel_err(hist.cpu().numpy(), z["rk4_hist"]) < HORIZON_TOL
    assert rel_err(best.cpu().numpy(), z["rk4_best"]) < HORIZON_TOL
    assert np.abs(U.cpu().numpy() - z["rk4_U_last"]).max() < 0.02 * lr + 1e-5
    _, g0, _ = ops.cost_grad(pk.handle, cu(x0), cu(U0), float(z["dt"]), 1, *ca, True, False)
    assert rel_err(g0.cpu().numpy(), z["rk4_grad0"]) < HORIZON_TOL
    U19, _, _ = ops.mpc_solve(pk.handle, cu(x0), cu(U0), float(z["dt"]), 1, *ca, lr, 0.9, 0.999, 1e-8, iters - 1, 0, False)
    _, g19, _ = ops.cost_grad(pk.handle, cu(x0), U19, float(z["dt"]), 1, *ca, True, False)
    assert rel_err(g19.cpu().numpy(), z["rk4_grad_last"]) < HORIZON_TOL


# ---------------------------------------------------------------------------------------------
# CTA pairs (option "tensor_pair", tcgen05 cta_group::2): two 128-instance tiles per cluster, one M = 256 MMA per product,
# each CTA staging half of every weight tile.  Same MMAs per row and the same element code as the single-CTA launch, so
# the results are required to be BIT-IDENTICAL to it for models without an R_net (for the others the two launches add the
# R_net sums in a different order: 2e-6) -- and, independently, within the stated bounds of the oracle.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cartpole_h256", "cartpole_h128", "canonical"])
@pytest.mark.parametrize("B,H,iters", [(256, 6, 3), (128 * 300, 4, 2), (128 * 22 - 5, 5, 2), (128 * 7, 5, 2)])
def test_cta_pair_bit_identical_to_single(tc_env, name, B, H, iters):
    """even tile counts run as CTA pairs (incl. more tiles than SMs and a ragged last tile); an odd tile count (7) falls
    back to the single-CTA launch; either way the solve equals the tensor_pair = 0 solve (bit for bit without an R_net)"""
    ops, get_tc = tc_env
    z, sd, pk = get_tc(name, 4)
    rng = np.random.default_rng(B + H)
    x0 = (rng.uniform(-1, 1, size=(B, 4)) * [1.0, 0.3, 0.5, 0.5]).astype(np.float32)
    U0 = rng.uniform(-2, 2, size=(B, H, 1)).astype(np.float32)
    _, _, ca = _cfg4_cost()
    outs = []
    try:
        for pair in (0, 1):
            pk.set_option("tensor_pair", pair)
            assert pk.get_option("tensor_pair") == pair
            U, hist, best = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, 0.015, 0.9, 0.999, 1e-8, iters, 1, True)
            torch.cuda.synchronize()
            outs.append((U.cpu().numpy(), hist.cpu().numpy(), best.cpu().numpy()))
    finally:
        pk.set_option("tensor_pair", 0)   # the default
    for a, b in zip(*outs):
        if KINDS[name] == "canonical":
            assert np.array_equal(a, b)           # same MMAs per row, same element code: bit for bit
        else:
            # models with an R_net: the single-CTA launch adds the R_net sums on its dedicated warps (one thread per
            # instance and half of the hidden units), the pair launch inside the element warps (per-lane partial sums):
            # same terms, different order of the FP32 additions
            assert rel_err(a, b) < 2e-6


def test_cta_pair_benchmarked_config_vs_oracle(tc_env):
    """bench.py's default job (hidden 256, H=50, RK4, 20 Adam iterations) on 256 of its instances through the CTA-pair
    launch, against the CPU oracle"""
    from oracle.phnn_oracle import OracleModel, set_threads
    import os
    ops, get_tc = tc_env
    z, sd, pk = get_tc("cartpole_h256", 4)
    set_threads(os.cpu_count() or 1)
    M = OracleModel(sd, "phnn")
    B, H, iters, lr = 256, 50, 20, 0.015
    x0 = _bench_inputs(B).numpy()
    U0 = np.zeros((B, H, 1), np.float32)
    Q, R, ca = _cfg4_cost()
    C = M.cost_struct(Q, R, np.zeros(4), -15.0, 15.0)
    Uo, histo, besto = M.mpc_solve(C, x0, U0, 0.02, "rk4", lr=lr, iters=iters)
    pk.set_option("tensor_pair", 1)
    try:
        U, hist, best = ops.mpc_solve(pk.handle, cu(x0), cu(U0), 0.02, 1, *ca, lr, 0.9, 0.999, 1e-8, iters, 0, True)
        torch.cuda.synchronize()
    finally:
        pk.set_option("tensor_pair", 0)
    assert rel_err(hist.cpu().numpy(), histo) < HORIZON_TOL
    assert rel_err(best.cpu().numpy(), besto) < HORIZON_TOL
    assert np.abs(U.cpu().numpy() - Uo).max() < 0.02 * lr + 1e-5


@pytest.mark.parametrize("mode", [4, 2])
@pytest.mark.parametrize("H", [100, 200])
def test_cfg5_shape_golden_reference(tc_env, H, mode):
    """BASELINE cfg5 horizons against the fixture recorded from the REFERENCE ITSELF (one hop, no oracle):
    tests/golden/cfg5_shape.npz = make_golden.gen_cfg5_shape (cfg4 weights, bench.py's first 16 instances, RK4, 4 Adam
    iterations; reference PyTorch autograd + torch.optim.Adam).  Stated long-horizon bounds of DESIGN.md section 2."""
    ops, get_tc = tc_env
    _, sd, pk = get_tc("cartpole_h256", mode)
    z, _ = load_golden("cfg5_shape")
    B, iters, lr = z["x0"].shape[0], int(z["iters"]), float(z["lr"])
    assert np.array_equal(_bench_inputs(B).numpy(), z["x0"])
    U0 = np.zeros((B, H, 1), np.float32)
    _, _, ca = _cfg4_cost()
    p = "h%d_" % H
    tol = HORIZON_TOL * H / 50.0
    U, hist, best = ops.mpc_solve(pk.handle, cu(z["x0"]), cu(U0), float(z["dt"]), 1, *ca, lr, 0.9, 0.999, 1e-8, iters, 0, True)
    assert rel_err(hist.cpu().numpy(), z[p + "hist"]) < tol
    assert rel_err(best.cpu().numpy(), z[p + "best"]) < tol
    assert np.abs(U.cpu().numpy() - z[p + "U_last"]).max() < 0.05 * lr
    _, g0, _ = ops.cost_grad(pk.handle, cu(z["x0"]), cu(U0), float(z["dt"]), 1, *ca, True, False)
    assert rel_err(g0.cpu().numpy(), z[p + "grad0"]) < 5 * tol
