"""Closed loop: plant step, driver bookkeeping and warm-start shift.

CPU part: the oracle's plant and the oracle composition (cast -> solve -> plant step [-> shift]) against the
closed loops recorded from the reference (CartPoleSimulator + MPCController / MPCControllerCanonical.control).
GPU part: the device-resident batched loop (phnn_mpc_b200.closed_loop) against the same golden trajectories."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.phnn_oracle import OracleModel, plant_step

CFG1 = dict(Q=[10.0, 200.0, 1.0, 10.0], R=[0.01], umin=-15.0, umax=15.0, lr=0.015, iters=30, H=20, mode="last")
CFG3 = dict(Q=[0.0, 1000.0, 0.0, 100.0], R=[1e-4], umin=-30.0, umax=30.0, lr=0.03, iters=50, H=10, mode="best")


def test_oracle_plant_matches_reference_bitwise():
    z, _ = load_golden("closed_loop")
    s = z["plant_traj"][0:1].copy()
    for k, u in enumerate(z["plant_u"]):
        s, d = plant_step(s, [u], 0.02)
        assert np.array_equal(s[0], z["plant_traj"][k + 1])          # float64 arithmetic in the same order
        assert bool(d[0]) == bool(z["plant_done"][k])


def oracle_closed_loop(M, cfg, x0s, steps, warm):
    C = M.cost_struct(cfg["Q"], cfg["R"], np.zeros(4), cfg["umin"], cfg["umax"])
    state = np.asarray(x0s, np.float64).copy()
    B = state.shape[0]
    traj, ctr = [state.copy()], []
    U0 = np.zeros((B, cfg["H"], 1), np.float32)
    for _ in range(steps):
        U, _, _ = M.mpc_solve(C, state.astype(np.float32), U0, 0.02, "euler", lr=cfg["lr"], iters=cfg["iters"],
                              return_mode=cfg["mode"])
        u = U[:, 0, 0]
        ctr.append(u.copy())
        state, _ = plant_step(state, u, 0.02)
        traj.append(state.copy())
        if warm:
            U0 = np.concatenate([U[:, 1:], np.zeros((B, 1, 1), np.float32)], 1)
    return np.stack(traj, 1), np.stack(ctr, 1)


@pytest.mark.parametrize("name,kind,cfg,warm", [("closed_loop", "phnn", CFG1, False),
                                                ("closed_loop_canonical", "canonical", CFG3, True)])
def test_oracle_closed_loop_matches_reference(name, kind, cfg, warm):
    z, sd = load_golden(name)
    M = OracleModel(sd, kind)
    steps = z["cl_u"].shape[1]
    traj, ctr = oracle_closed_loop(M, cfg, z["cl_x0"], steps, warm)
    assert np.abs(ctr - z["cl_u"]).max() < 0.05 * cfg["lr"]
    assert np.abs(traj - z["cl_traj"]).max() < 1e-5


def _batched(name, kind, cfg):
    from phnn_mpc_b200.batched import BatchedMPC, CostSpec
    from phnn_mpc_b200.packing import PackedModel
    z, sd = load_golden(name)
    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind)
    spec = CostSpec.make(4, 1, cfg["Q"], cfg["R"], None, cfg["umin"], cfg["umax"])
    return z, sd, BatchedMPC(pk, cfg["H"], 0.02, spec, integrator="euler", lr=cfg["lr"], iters=cfg["iters"],
                             return_mode=cfg["mode"])


@pytest.mark.gpu
def test_device_closed_loop_phnn_cold_start():
    from phnn_mpc_b200.closed_loop import ClosedLoopBatch
    z, sd, mpc = _batched("closed_loop", "phnn", CFG1)
    steps = z["cl_u"].shape[1]
    loop = ClosedLoopBatch(mpc, warm_start=False, target=np.zeros(4), tolerance=[0.1, 0.1, 0.05, 0.05], min_duration=0.04)
    out = loop.run(z["cl_x0"], steps)
    assert np.abs(out["controls"].cpu().numpy() - z["cl_u"]).max() < 0.05 * CFG1["lr"]
    assert np.abs(out["states"].cpu().numpy() - z["cl_traj"]).max() < 1e-5
    assert np.abs(out["energies"].cpu().numpy() - z["cl_H"]).max() < 1e-5 * max(1.0, np.abs(z["cl_H"]).max())
    assert (out["done_step"].cpu().numpy() == -1).all()
    # stability bookkeeping: an instance whose recorded states stay within tolerance for >= 2 consecutive steps
    within = z["cl_within"]
    run2 = np.array([any(within[b, k] and within[b, k + 1] for k in range(steps - 1)) for b in range(within.shape[0])])
    assert np.array_equal(out["stability_achieved"].cpu().numpy(), run2)


@pytest.mark.gpu
def test_device_closed_loop_canonical_warm_start():
    from phnn_mpc_b200.closed_loop import ClosedLoopBatch
    z, sd, mpc = _batched("closed_loop_canonical", "canonical", CFG3)
    steps = z["cl_u"].shape[1]
    out = ClosedLoopBatch(mpc, warm_start=True, log_energy=False).run(z["cl_x0"], steps)
    assert np.abs(out["controls"].cpu().numpy() - z["cl_u"]).max() < 0.05 * CFG3["lr"]
    assert np.abs(out["states"].cpu().numpy() - z["cl_traj"]).max() < 1e-5


@pytest.mark.gpu
def test_device_closed_loop_batch_termination_and_oracle():
    """200 plants, some starting beyond the 0.5 rad limit region: terminated plants freeze, the rest follow the oracle."""
    from phnn_mpc_b200.closed_loop import ClosedLoopBatch
    z, sd, mpc = _batched("closed_loop", "phnn", dict(CFG1, iters=5, H=8))
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-1, 1, size=(200, 4)) * [1.0, 0.3, 0.5, 0.5]
    x0[:10, 1] = 0.49
    x0[:10, 3] = 3.0                                         # falls over within a step or two
    steps = 4
    out = ClosedLoopBatch(mpc, warm_start=False).run(x0, steps)
    M = OracleModel(sd, "phnn")
    traj, ctr = oracle_closed_loop(M, dict(CFG1, iters=5, H=8), x0, steps, False)
    done = out["done_step"].cpu().numpy()
    assert (done[:10] > 0).all() and (done[10:] == -1).sum() > 150
    st = out["states"].cpu().numpy()
    alive = done == -1
    assert np.abs(st[alive] - traj[alive]).max() < 1e-5
    assert np.abs(out["controls"].cpu().numpy()[alive] - ctr[alive]).max() < 0.05 * CFG1["lr"]
    for b in range(10):                                       # frozen after termination
        k = done[b]
        assert np.array_equal(st[b, k], st[b, -1]) and np.all(out["controls"].cpu().numpy()[b, k:] == 0)


@pytest.mark.gpu
def test_shift_controls_kernel():
    import ctypes
    from phnn_mpc_b200 import _lib
    U = torch.arange(3 * 5, dtype=torch.float32, device="cuda").reshape(3, 5, 1)
    out = torch.empty_like(U)
    _lib.check(_lib.lib().phnn_shift_controls(U.data_ptr(), out.data_ptr(), 3, 5,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "shift")
    ref = torch.cat([U[:, 1:], torch.zeros(3, 1, 1, device="cuda")], 1)
    assert torch.equal(out, ref)
    assert _lib.lib().phnn_shift_controls(U.data_ptr(), U.data_ptr(), 3, 5, None) == _lib.E_ARG
