"""The reference's own driver code, UNCHANGED, against the drop-in modules (SURVEY.md section 4 "integration" tier,
section 8b): `run_mpc_control` of scripts/run_cartpole_mpc.py:91-182 and `simulate_mpc_control` of
scripts/run_mpc_canonical.py:26-104 are imported from the byte-for-byte copies staged under oracle/_ref
(oracle/fetch_ref.py) with `phnn_mpc_b200/dropin` and a no-op matplotlib stub AHEAD of the scripts' own
`sys.path.append('src')`, and are driven with the reference's own YAML files.  The closed loops they produce must match
the ones recorded from the reference itself (tests/golden/closed_loop*.npz, made by make_golden.gen_closed_loop)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden

REF = os.path.join(REPO, "oracle", "_ref")
DROPIN = os.path.join(REPO, "phnn_mpc_b200", "dropin")
STUBS = os.path.join(REPO, "tests", "stubs")
REF_MODULES = ("pHNN", "pHNN_canonical", "NN", "mass_matrix", "coordinate_transforms", "integrators", "mpc_controller",
               "mpc_controller_canonical", "cartpole_simulator")


def _load_driver(name):
    """import oracle/_ref/scripts/<name>.py the way `python scripts/<name>.py` would resolve its imports when
    PYTHONPATH holds the drop-in directory, with cwd = the reference root"""
    if not os.path.exists(os.path.join(REF, "MANIFEST.json")):
        pytest.skip("oracle/_ref is not staged (run python oracle/fetch_ref.py where /root/reference exists)")
    for m in REF_MODULES:
        sys.modules.pop(m, None)
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    os.chdir(REF)
    sys.path[:0] = [STUBS, DROPIN]
    try:
        spec = importlib.util.spec_from_file_location("_ref_driver_" + name, os.path.join(REF, "scripts", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)      # runs the script's own imports: sys.path.append('src'); from pHNN import pHNN ...
    finally:
        os.chdir(saved_cwd)
    return mod, saved_path


def _restore(saved_path):
    sys.path[:] = saved_path
    for m in REF_MODULES:
        sys.modules.pop(m, None)


def test_driver_imports_resolve_to_dropins_and_reference_plant():
    """no GPU needed: the unchanged scripts bind the drop-in model / controller classes and the reference's own plant"""
    mod, saved = _load_driver("run_cartpole_mpc")
    try:
        assert os.path.abspath(sys.modules[mod.pHNN.__module__].__file__).startswith(DROPIN)
        assert os.path.abspath(sys.modules[mod.MPCController.__module__].__file__).startswith(DROPIN)
        assert os.path.abspath(sys.modules[mod.CartPoleSimulator.__module__].__file__).startswith(os.path.join(REF, "src"))
    finally:
        _restore(saved)
    mod, saved = _load_driver("run_mpc_canonical")
    try:
        assert os.path.abspath(sys.modules[mod.pHNN_Canonical.__module__].__file__).startswith(DROPIN)
        assert os.path.abspath(sys.modules[mod.create_mpc_controller.__module__].__file__).startswith(DROPIN)
        cfg = mod.yaml.safe_load(open(os.path.join(REF, "pole_stabilization_config.yaml")))
        assert cfg["mpc"]["horizon"] == 10 and cfg["mpc"]["optimizer_steps"] == 50
    finally:
        _restore(saved)


@pytest.mark.gpu
def test_unchanged_run_cartpole_mpc_closed_loop():
    """scripts/run_cartpole_mpc.py: load_config + create_mpc_from_config + run_mpc_control, reference YAML, 5+1 steps"""
    z, sd = load_golden("closed_loop")
    mod, saved = _load_driver("run_cartpole_mpc")
    try:
        cfg_path = os.path.join(REF, "cartpole_mpc_config.yaml")
        config = mod.load_config(cfg_path)
        torch.manual_seed(0)
        model = mod.pHNN(cfg_path)                       # drop-in class, reference YAML, reference seeding order
        for k, v in model.state_dict().items():
            assert np.array_equal(v.numpy(), sd[k]), k   # same random init as the reference model of the fixture
        model.eval()
        ctrl = mod.create_mpc_from_config(model, config)
        assert (ctrl.horizon, ctrl.max_iterations, ctrl.u_min, ctrl.u_max) == (20, 30, -15.0, 15.0)
        steps = z["cl_u"].shape[1]
        for b in range(z["cl_x0"].shape[0]):
            sim = mod.CartPoleSimulator(dt=config["cartpole"]["dt"])
            states, controls, hams, achieved, dur = mod.run_mpc_control(sim, ctrl, z["cl_x0"][b].copy(), steps, config,
                                                                        verbose=False)
            assert states.shape == (steps + 1, 4) and controls.shape == (steps,)
            assert np.abs(controls - z["cl_u"][b]).max() < 0.05 * 0.015
            assert np.abs(states - z["cl_traj"][b]).max() < 1e-5
            assert np.abs(hams - z["cl_H"][b]).max() < 1e-5 * max(1.0, np.abs(z["cl_H"]).max())
    finally:
        _restore(saved)


@pytest.mark.gpu
def test_unchanged_run_mpc_canonical_closed_loop():
    """scripts/run_mpc_canonical.py: simulate_mpc_control with the canonical drop-in pair and the reference's
    pole_stabilization_config.yaml (warm start from the shifted previous plan)"""
    z, sd = load_golden("closed_loop_canonical")
    mod, saved = _load_driver("run_mpc_canonical")
    try:
        torch.manual_seed(0)
        model = mod.pHNN_Canonical(os.path.join(REF, "cartpole_mpc_config.yaml"))
        for k, v in model.state_dict().items():
            assert np.array_equal(v.numpy(), sd[k]), k
        model.eval()
        pcfg = mod.yaml.safe_load(open(os.path.join(REF, "pole_stabilization_config.yaml")))
        ctrl = mod.create_mpc_controller(model, pcfg)
        assert (ctrl.horizon, ctrl.optimizer_steps) == (10, 50)
        steps = z["cl_u"].shape[1]
        for b in range(z["cl_x0"].shape[0]):
            sim = mod.CartPoleSimulator(pcfg["cartpole"]["dt"])
            states, controls, costs, solve_times = mod.simulate_mpc_control(sim, ctrl, z["cl_x0"][b].copy(), num_steps=steps,
                                                                            verbose=False)
            assert states.shape == (steps + 1, 4) and controls.shape == (steps, 1) and costs.shape == (steps,)
            assert np.abs(controls[:, 0] - z["cl_u"][b]).max() < 0.05 * 0.03
            assert np.abs(states - z["cl_traj"][b]).max() < 1e-5
            assert (solve_times > 0).all()
    finally:
        _restore(saved)
