"""Forward-only tcgen05 instantiation for the n = 2 pHNN (pendulum model of BASELINE cfg2: h = 64, learned G,
src/pHNN.py:52-100 with the G_net branch :86-92, rollouts of src/integrators.py:128-258) against the golden
rollouts recorded from the reference, the CPU oracle and the latency kernel.  Everything goes through the C ABI.

Tolerances: 1e-5 per evaluation, 1e-4 per 10-step horizon, 2e-4 on the 100-step cfg2 rollout (as for the other
kernels), and the FP64 tie-breaker with factor 4 instead of 2 (floor 1e-4): the tensor-core operands carry 22
significand bits (FP16 hi + lo) instead of 24, so a per-evaluation error of up to four FP32 roundings is the
stated bound of this path."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


@pytest.fixture(scope="module")
def pend():
    from phnn_mpc_b200 import ops
    from phnn_mpc_b200.packing import PackedModel
    z, sd = load_golden("pendulum")
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn")
    assert pk.get_option("tensor_mode") == 4 and pk.get_option("tensor_fwd_min_batch") > 0
    pk.set_option("tensor_fwd_min_batch", 1)   # every forward job of this pack on the tcgen05 kernel
    return ops, z, sd, pk


def test_forward_golden_and_ragged(pend):
    from oracle.phnn_oracle import OracleModel
    ops, z, sd, pk = pend
    dx, H = ops.forward(pk.handle, cu(z["rand_x"]), cu(z["rand_u"]))
    assert rel_err(dx.cpu().numpy(), z["rand_dx"]) < 1e-5
    assert rel_err(H.cpu().numpy(), z["rand_H"]) < 1e-5
    M = OracleModel(sd, "phnn")
    for B in (1, 127, 129, 300):   # partial tiles
        rng = np.random.default_rng(B)
        x = rng.normal(size=(B, 2)).astype(np.float32) * 2
        u = rng.normal(size=(B, 1)).astype(np.float32) * 3
        dx, H = ops.forward(pk.handle, cu(x), cu(u))
        dxo, Ho = M.forward(x, u)
        assert rel_err(dx.cpu().numpy(), dxo) < 1e-5
        assert rel_err(H.cpu().numpy(), Ho) < 1e-5


def test_anchor_rollouts_both_energy_orderings(pend):
    ops, z, sd, pk = pend
    U10 = np.repeat(z["anchor_u"][:, None, :], 10, 1)
    for integ, iid in (("rk4", 1), ("euler", 0)):
        tr, en = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, iid, 1)
        assert rel_err(tr.cpu().numpy(), z["anchor_traj_" + integ]) < 1e-4
        assert rel_err(en.cpu().numpy(), z["anchor_en_" + integ]) < 1e-4
        tr2, en2 = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, iid, 2)
        assert rel_err(tr2.cpu().numpy(), z["anchor_traj2_" + integ]) < 1e-4
        assert rel_err(en2.cpu().numpy(), z["anchor_en2_" + integ]) < 1e-4
    tr, _ = ops.rollout(pk.handle, cu(z["anchor_x"]), cu(U10), 0.05, 1, 0)
    np.testing.assert_allclose(tr[:, -1].cpu().numpy(), [[0.300435662, -2.837836504], [-0.848876774, 4.260723591]], rtol=2e-5)


@pytest.mark.parametrize("integ", ["rk4", "euler"])
def test_cfg2_rollout(pend, integ):
    from oracle.phnn_oracle import OracleModel
    ops, z, sd, pk = pend
    tr, en = ops.rollout(pk.handle, cu(z["cfg2_x0"]), cu(z["cfg2_U"]), 0.05, {"euler": 0, "rk4": 1}[integ], 1)
    tr, en = tr.cpu().numpy(), en.cpu().numpy()
    assert rel_err(tr, z["cfg2_traj_" + integ]) < 2e-4
    assert rel_err(en, z["cfg2_en_" + integ]) < 2e-4
    o64 = OracleModel(sd, "phnn", np.float64).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    o32 = OracleModel(sd, "phnn", np.float32).rollout(z["cfg2_x0"], z["cfg2_U"], 0.05, integ)
    assert rel_err(tr, o64) < max(4 * rel_err(o32, o64), 1e-4)
    assert rel_err(tr[:, :11], z["cfg2_traj_" + integ][:, :11]) < 1e-5


def test_routes_agree_and_cost_without_gradient(pend):
    """same inputs through the latency kernel and the tcgen05 kernel (10 steps: before rounding is amplified), and the
    cost-only job against the oracle"""
    from oracle.phnn_oracle import OracleModel
    ops, z, sd, pk = pend
    rng = np.random.default_rng(5)
    B, T = 700, 10
    x0 = (rng.uniform(-1, 1, size=(B, 2)) * [2.0, 2.0]).astype(np.float32)
    U = rng.uniform(-1, 1, size=(B, T, 1)).astype(np.float32)
    tr_tc, en_tc = ops.rollout(pk.handle, cu(x0), cu(U), 0.05, 1, 1)
    pk.set_option("tensor_fwd_min_batch", 0)
    try:
        tr_lat, en_lat = ops.rollout(pk.handle, cu(x0), cu(U), 0.05, 1, 1)
    finally:
        pk.set_option("tensor_fwd_min_batch", 1)
    assert rel_err(tr_tc.cpu().numpy(), tr_lat.cpu().numpy()) < 1e-5
    assert rel_err(en_tc.cpu().numpy(), en_lat.cpu().numpy()) < 1e-5
    Q = torch.diag(torch.tensor([5.0, 1.0]))
    args = (Q, torch.tensor([[0.1]]), torch.tensor([0.5, 0.0]), True, -2.0, 2.0, None, None, 1000.0)
    cost, g, tr = ops.cost_grad(pk.handle, cu(x0), cu(U), 0.05, 1, *args, False, True)
    M = OracleModel(sd, "phnn")
    C = M.cost_struct([5.0, 1.0], [0.1], np.array([0.5, 0.0]), -2.0, 2.0)
    co, tro = M.cost_grad(C, x0, U, 0.05, "rk4", want_grad=False, want_traj=True)
    assert rel_err(cost.cpu().numpy(), co) < 1e-4
    assert rel_err(tr.cpu().numpy(), tro) < 1e-4


def test_fixed_G_pendulum_shape():
    """(pHNN, fixed G, n = 2, h = 64) random model against the oracle on the tcgen05 route"""
    from oracle.phnn_oracle import OracleModel
    from phnn_mpc_b200 import ops
    from phnn_mpc_b200.packing import PackedModel
    g = torch.Generator().manual_seed(11)
    h = 64
    def lin(o, i, s=1.0):
        return (torch.randn(o, i, generator=g) * s / np.sqrt(i)).float(), (torch.randn(o, generator=g) * 0.1).float()
    sd = {}
    for net, dims in (("H_net", (2, h, h, 1)), ("R_net", (2, h, 4))):
        for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
            w, b = lin(o, i, 1.5)
            sd["%s.net.%d.weight" % (net, 2 * li)] = w
            sd["%s.net.%d.bias" % (net, 2 * li)] = b
    sd["J"] = torch.tensor([[0.0, 1.0], [-0.3, 0.0]])
    sd["G_fixed"] = torch.tensor([[0.0], [1.0]])
    pk = PackedModel(sd, "phnn")
    pk.set_option("tensor_fwd_min_batch", 1)
    sdn = {k: v.numpy() for k, v in sd.items()}
    M = OracleModel(sdn, "phnn")
    rng = np.random.default_rng(2)
    B, T = 257, 20
    x0 = rng.normal(size=(B, 2)).astype(np.float32)
    U = rng.normal(size=(B, T, 1)).astype(np.float32)
    dx, H = ops.forward(pk.handle, cu(x0), cu(U[:, 0]))
    dxo, Ho = M.forward(x0, U[:, 0])
    assert rel_err(dx.cpu().numpy(), dxo) < 1e-5
    assert rel_err(H.cpu().numpy(), Ho) < 1e-5
    tr, _ = ops.rollout(pk.handle, cu(x0), cu(U), 0.02, 1, 0)
    assert rel_err(tr.cpu().numpy(), M.rollout(x0, U, 0.02, "rk4")) < 1e-4


def test_sparse_and_dense_tiles_agree(pend):
    """jobs of up to 64 instances per SM run on 64-instance tiles (two warps of a TMEM quadrant split the K-blocks and add
    their sums through shared memory), larger ones on 128-instance tiles: same results within one evaluation's rounding,
    ragged last tiles included; a job beyond the 64-per-SM limit against the oracle"""
    from oracle.phnn_oracle import OracleModel
    ops, z, sd, pk = pend
    assert pk.get_option("tensor_fwd_sparse") == 1
    rng = np.random.default_rng(9)
    M = OracleModel(sd, "phnn")
    for B in (1, 63, 65, 700):
        x0 = (rng.uniform(-1, 1, size=(B, 2)) * [2.0, 2.0]).astype(np.float32)
        U = rng.uniform(-1, 1, size=(B, 10, 1)).astype(np.float32)
        tr_s, en_s = ops.rollout(pk.handle, cu(x0), cu(U), 0.05, 1, 2)
        pk.set_option("tensor_fwd_sparse", 0)
        try:
            tr_d, en_d = ops.rollout(pk.handle, cu(x0), cu(U), 0.05, 1, 2)
        finally:
            pk.set_option("tensor_fwd_sparse", 1)
        assert rel_err(tr_s.cpu().numpy(), tr_d.cpu().numpy()) < 1e-5
        assert rel_err(en_s.cpu().numpy(), en_d.cpu().numpy()) < 1e-5
        tro, eno = M.rollout(x0, U, 0.05, "rk4", energy_mode=2)
        assert rel_err(tr_s.cpu().numpy(), tro) < 1e-4
        assert rel_err(en_s.cpu().numpy(), eno) < 1e-4
    B = 64 * 148 + 100   # one more than fits 64 per SM on a B200: 128-instance tiles
    x0 = (rng.uniform(-1, 1, size=(B, 2)) * [2.0, 2.0]).astype(np.float32)
    U = rng.uniform(-1, 1, size=(B, 10, 1)).astype(np.float32)
    tr, _ = ops.rollout(pk.handle, cu(x0), cu(U), 0.05, 1, 0)
    sel = np.r_[0:40, B - 40:B]
    assert rel_err(tr.cpu().numpy()[sel], M.rollout(x0[sel], U[sel], 0.05, "rk4")) < 1e-4
