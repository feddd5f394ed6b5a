"""CPU-only tests: host logic, the C-ABI library's exports, drop-in construction parity and the
multi-process sharding path (gloo, world_size 2).  No CUDA compute is invoked here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import CONFIGS, REPO, load_golden


def test_library_exports_every_declared_symbol():
    """libphnn_mpc.so loads and exports what include/phnn_mpc.h declares."""
    from phnn_mpc_b200.build import build_library
    lib = ctypes.CDLL(build_library())
    header = open(os.path.join(REPO, "include", "phnn_mpc.h")).read()
    declared = set(re.findall(r"\b(phnn_[a-z0-9_]+)\s*\(", header))
    declared -= {"phnn_model_desc", "phnn_cost_desc", "phnn_pack"}
    assert len(declared) >= 11
    for sym in declared:
        assert hasattr(lib, sym), sym
    from phnn_mpc_b200 import _lib
    assert set(_lib.EXPORTS) <= declared
    lib.phnn_version.restype = ctypes.c_int
    assert lib.phnn_version() >= 100


def test_argument_errors_without_gpu():
    """bad arguments are rejected before any CUDA call; messages come from phnn_last_error()."""
    from phnn_mpc_b200 import _lib
    L = _lib.lib()
    assert L.phnn_pack_create(None, 0, None) == _lib.E_ARG
    assert b"null" in L.phnn_last_error()
    d = _lib.ModelDesc()
    d.m = 2
    h = ctypes.c_void_p()
    assert L.phnn_pack_create(ctypes.byref(d), 0, ctypes.byref(h)) == _lib.E_UNSUPPORTED
    assert L.phnn_workspace_bytes(None, 10, 10, 0) == 0
    with pytest.raises(RuntimeError):
        _lib.check(_lib.E_UNSUPPORTED, "x")
    with pytest.raises(ValueError):
        _lib.check(_lib.E_INTEGRATOR, "x")


def test_ops_refuse_cpu_tensors():
    """no CPU fallback: the ops raise on host tensors instead of computing something else"""
    from phnn_mpc_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.forward(0, torch.zeros(2, 4), torch.zeros(2, 1))
    with pytest.raises(ValueError, match="Unknown integrator"):
        ops.integrator_id("midpoint")


def test_dropin_init_matches_reference_seeding():
    """torch.manual_seed(s); pHNN(cfg) yields the reference's weights bit for bit (the golden
    fixtures store the state_dicts the reference produced with the same seeds)."""
    from phnn_mpc_b200.dropin.pHNN import pHNN
    from phnn_mpc_b200.dropin.pHNN_canonical import pHNN_Canonical
    torch.manual_seed(0)
    m = pHNN(os.path.join(CONFIGS, "cartpole_phnn.yaml"))
    _, sd = load_golden("cartpole_h128")
    assert sorted(m.state_dict()) == sorted(sd)
    for k, v in sd.items():
        assert np.array_equal(m.state_dict()[k].numpy(), v), k
    torch.manual_seed(0)
    c = pHNN_Canonical(os.path.join(CONFIGS, "cartpole_phnn.yaml"))
    _, sdc = load_golden("canonical")
    assert sorted(c.state_dict()) == sorted(sdc)
    for k in ("H_net.net.0.weight", "H_net.net.2.weight", "H_net.net.4.bias", "J", "G"):
        assert np.array_equal(c.state_dict()[k].numpy(), sdc[k]), k
    c.load_state_dict({k: torch.from_numpy(v) for k, v in sdc.items()})      # reference checkpoints load
    p = pHNN(os.path.join(CONFIGS, "pendulum_phnn.yaml"))
    _, sdp = load_golden("pendulum")
    p.load_state_dict({k: torch.from_numpy(v) for k, v in sdp.items()})
    assert p.G_net is not None and not hasattr(p, "G_fixed")


def test_dropin_constant_mass_matrix(tmp_path):
    """mass_matrix.type 'constant' (MassMatrixNetwork, src/pHNN_canonical.py:79-86, src/mass_matrix.py:15-216): the drop-in
    builds the same module tree (state_dict keys of the fixture recorded from the reference), loads the reference's
    state_dict, and its (a, b, c) are the reference's M = L L^T; the configuration-dependent types raise."""
    import yaml
    from phnn_mpc_b200.dropin.pHNN_canonical import pHNN_Canonical
    from phnn_mpc_b200.packing import constant_mass_abc
    cfg = yaml.safe_load(open(os.path.join(CONFIGS, "cartpole_phnn.yaml")))
    cfg["model"]["mass_matrix"] = {"type": "constant", "init_scale": 1.0}
    path = tmp_path / "constM.yaml"
    path.write_text(yaml.safe_dump(cfg))
    torch.manual_seed(0)
    c = pHNN_Canonical(str(path))
    z, sd = load_golden("canonical_constM")
    assert sorted(c.state_dict()) == sorted(sd)
    for k in ("H_net.net.0.weight", "H_net.net.2.weight", "H_net.net.4.bias", "J", "G"):   # same RNG order as the reference
        assert np.array_equal(c.state_dict()[k].numpy(), sd[k]), k
    c.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    a, b, cc = constant_mass_abc(sd["M_net.L_tril"])
    np.testing.assert_allclose([[a, b], [b, cc]], z["M"], rtol=1e-6)
    q = torch.zeros(3, 2)
    np.testing.assert_allclose(c.M_net(q)[0].detach().numpy(), z["M"], rtol=1e-6)
    np.testing.assert_allclose((c.M_net(q)[1] @ c.M_net.inverse(q)[1]).detach().numpy(), np.eye(2), atol=1e-5)
    for t in ("diagonal", "full"):
        cfg["model"]["mass_matrix"] = {"type": t}
        path.write_text(yaml.safe_dump(cfg))
        with pytest.raises(NotImplementedError):
            pHNN_Canonical(str(path))


def test_dropin_dropout_mlps_eval_mode(tmp_path):
    """MLPs configured with dropout > 0 (src/NN.py:16-25): the drop-in builds the reference's module tree (Linear layers at
    net.0 / net.3 / net.6, same RNG order), loads its state_dict, maps the keys to the dropout-free layout for the kernels
    and refuses a forward in training mode (Dropout is the identity only in eval mode, src/mpc_controller.py:44)."""
    import yaml
    from phnn_mpc_b200.dropin.pHNN import pHNN
    from phnn_mpc_b200.packing import normalize_state_dict
    cfg = yaml.safe_load(open(os.path.join(CONFIGS, "cartpole_phnn.yaml")))
    cfg["model"]["H_mlp"]["dropout"] = 0.1
    cfg["model"]["R_mlp"]["dropout"] = 0.1
    path = tmp_path / "dropout.yaml"
    path.write_text(yaml.safe_dump(cfg))
    torch.manual_seed(21)
    m = pHNN(str(path))
    z, sd = load_golden("cartpole_h128_dropout")
    assert sorted(m.state_dict()) == sorted(sd)
    for k, v in sd.items():
        assert np.array_equal(m.state_dict()[k].numpy(), v), k
    nsd = normalize_state_dict(sd)
    assert "H_net.net.4.weight" in nsd and "H_net.net.6.weight" not in nsd and "R_net.net.2.bias" in nsd
    assert np.array_equal(nsd["H_net.net.2.weight"], sd["H_net.net.3.weight"])
    assert normalize_state_dict(load_golden("cartpole_h128")[1]).keys() == load_golden("cartpole_h128")[1].keys()
    m.train()
    with pytest.raises(RuntimeError, match="eval"):
        m(torch.zeros(1, 4), torch.zeros(1, 1))
    with pytest.raises(NotImplementedError):
        normalize_state_dict({"H_net.net.0.weight": np.zeros((8, 4)), "H_net.net.1.weight": np.zeros(8)})   # LayerNorm
    # bias: false (src/NN.py:19,25): same module tree as the reference, zero biases for the kernels
    from phnn_mpc_b200.dropin.pHNN_canonical import pHNN_Canonical
    cfg = yaml.safe_load(open(os.path.join(CONFIGS, "cartpole_phnn.yaml")))
    cfg["model"]["H_mlp"]["bias"] = False
    path.write_text(yaml.safe_dump(cfg))
    torch.manual_seed(33)
    c = pHNN_Canonical(str(path))
    _, sdn = load_golden("canonical_nobias")
    assert sorted(c.state_dict()) == sorted(sdn) and "H_net.net.0.bias" not in sdn
    assert np.array_equal(c.state_dict()["H_net.net.2.weight"].numpy(), sdn["H_net.net.2.weight"])
    nb = normalize_state_dict(sdn)
    assert all(np.all(nb["H_net.net.%d.bias" % i] == 0) for i in (0, 2, 4)) and nb["H_net.net.4.bias"].shape == (1,)


def test_dropin_surface_matches_reference_signatures():
    import inspect
    from phnn_mpc_b200.dropin import integrators, mpc_controller, mpc_controller_canonical
    sig = inspect.signature(mpc_controller.MPCController.__init__)
    assert list(sig.parameters)[1:] == ["phnn_model", "horizon", "dt", "Q", "R", "target_state", "u_min", "u_max",
                                        "x_min", "x_max", "optimizer_type", "lr", "max_iterations"]
    assert sig.parameters["lr"].default == 0.1 and sig.parameters["max_iterations"].default == 50
    sig = inspect.signature(mpc_controller_canonical.MPCControllerCanonical.__init__)
    assert list(sig.parameters)[1:] == ["model", "horizon", "dt", "Q", "R", "x_target", "u_min", "u_max",
                                        "optimizer_steps", "learning_rate", "verbose"]
    assert sig.parameters["u_min"].default == -10.0 and sig.parameters["horizon"].default == 20
    for fn in ("euler_step", "rk4_step", "rk4_step_with_energy", "rollout_trajectory",
               "rollout_trajectory_differentiable"):
        assert callable(getattr(integrators, fn))
    assert list(inspect.signature(integrators.rollout_trajectory_differentiable).parameters) == [
        "model", "y0", "controls", "dt", "integrator", "return_energies"]
    assert callable(mpc_controller.create_mpc_from_config) and callable(mpc_controller_canonical.create_mpc_controller)


def test_controller_construction_and_host_cost():
    from phnn_mpc_b200.dropin.pHNN import pHNN
    from phnn_mpc_b200.dropin.mpc_controller import MPCController
    torch.manual_seed(0)
    m = pHNN(os.path.join(CONFIGS, "cartpole_phnn.yaml"))
    c = MPCController(m, 20, 0.02, [10.0, 200.0, 1.0, 10.0], 0.01, [0, 0, 0, 0], -15.0, 15.0, lr=0.015,
                      max_iterations=30)
    assert c.horizon == 20 and c.u_min == -15.0 and c.max_iterations == 30 and c.model is m
    states = torch.arange(12, dtype=torch.float32).reshape(3, 4) / 10
    controls = torch.tensor([[1.0], [-2.0]])
    expect = sum(float(s @ c.Q @ s) for s in states) + 0.01 * 5.0
    assert abs(float(c.compute_cost(states, controls)) - expect) < 1e-4
    bad = MPCController(m, 5, 0.02, [1.0] * 4, 0.1, optimizer_type="SGD")
    with pytest.raises(ValueError):
        bad._check_optimizer()
    spec = c._spec()
    assert spec.op_args()[3] is True and spec.Q.shape == (4, 4)
    half = MPCController(m, 5, 0.02, [1.0] * 4, 0.1, u_min=-1.0)      # one bound only: no clamp (reference :180)
    assert half._spec().op_args()[3] is False


def test_shard_bounds_cover_batch():
    from phnn_mpc_b200.distributed import shard_bounds
    for B in (0, 1, 7, 64, 1000, 1048576):
        for ws in (1, 2, 3, 8):
            cuts = [shard_bounds(B, ws, r) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from phnn_mpc_b200.distributed import sharded_solve, shard_bounds
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
B = 37
x0 = torch.arange(B * 4, dtype=torch.float32).reshape(B, 4)
U0 = torch.arange(B * 5, dtype=torch.float32).reshape(B, 5, 1)
ITERS = 18                 # equals the size of rank 1's shard (37 = 19 + 18): the old shape test gathered cost_hist on
                           # one rank only and along the wrong axis (ADVICE r1)
def fake_solve(x, U):      # stands in for BatchedMPC.solve: any per-instance map
    hist = torch.arange(ITERS, dtype=torch.float32)[:, None] * 100 + x[:, 0][None, :]
    return {"U": U * 2 + x[:, :1, None], "u0": (U * 2 + x[:, :1, None])[:, 0], "best_cost": x.sum(1), "cost_hist": hist}
out = sharded_solve(fake_solve, x0, U0)
ref = fake_solve(x0, U0)
ok = all(torch.equal(out[k], ref[k]) for k in ("U", "u0", "best_cost", "cost_hist")) and out["cost_hist"].shape == (ITERS, B)
out2 = sharded_solve(lambda x, U: dict(fake_solve(x, U), cost_hist=None), x0, U0)
ok = ok and "cost_hist" not in out2 and torch.equal(out2["U"], ref["U"])
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_sharded_solve_gloo_world2(tmp_path):
    """N>1 path on CPU: two gloo ranks, contiguous shards, final all_gather equals the unsharded result."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", str(script), REPO]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("OK") == 2


def test_packed_weight_file_roundtrip_and_reference_checkpoint_layouts(tmp_path):
    """PHNNPK01 file <-> state_dict, and both .pth layouts the reference's drivers accept."""
    from phnn_mpc_b200 import weights_io
    from phnn_mpc_b200.dropin.pHNN import pHNN
    for name in ("pendulum", "cartpole_h128", "canonical"):
        _, sd = load_golden(name)
        f = weights_io.save_packed(str(tmp_path / (name + ".phnnpk")), sd)
        back, meta = weights_io.load_packed(f)
        assert sorted(back) == sorted(sd) and all(np.array_equal(back[k], sd[k]) for k in sd)
        assert meta["kind"] == ("canonical" if name == "canonical" else "phnn")
        assert meta["learned_G"] == (name == "pendulum") and meta["h"] == sd["H_net.net.0.weight"].shape[0]
    _, sd = load_golden("cartpole_h128")
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    raw, wrapped = str(tmp_path / "raw.pth"), str(tmp_path / "wrapped.pth")
    torch.save(tsd, raw)
    torch.save({"model_state_dict": tsd, "epoch": 860}, wrapped)
    for p in (raw, wrapped):
        got = weights_io.load_reference_checkpoint(p)
        assert all(np.array_equal(got[k], sd[k]) for k in sd)
    m = pHNN(os.path.join(CONFIGS, "cartpole_phnn.yaml"))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights_io.load_packed(
        weights_io.save_packed(str(tmp_path / "x.phnnpk"), sd))[0].items()})
    with pytest.raises(ValueError):
        (tmp_path / "bad.bin").write_bytes(b"notapack" + b"\0" * 64)
        weights_io.load_packed(str(tmp_path / "bad.bin"))


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm: the unmodified reference PyTorch path staged under oracle/_ref, on the
    host cores) prints ONE JSON line with the keys the driver reads and the SAME config / steps / warmup as our arm
    would report; no GPU, no kernel of ours on that path."""
    import json
    import subprocess
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    from oracle import fetch_ref
    if os.path.isdir(fetch_ref.DEFAULT_REF):
        fetch_ref.fetch(quiet=True)
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "small", "--ref-budget", "12"], capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cartpole_mpc_solves_per_s" and d["unit"] == "solves/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    import bench
    cfg, _, _ = bench.workload_config("small", list(bench.WORKLOADS["small"]), 1)
    assert d["config"] == cfg                      # what our arm prints for the same command line
    if fetch_ref.available():
        assert d["cpu_baseline"]["kind"] == "reference-pytorch" and d["cpu_baseline"]["torch_threads"] >= 1
        assert "oracle/_ref" in d["cpu_baseline"]["sample"]
    else:
        assert d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d.get("gpu_launches", 0) == 0
