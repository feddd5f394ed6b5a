"""oracle/ref_torch.py -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the UNMODIFIED reference PyTorch implementation staged under oracle/_ref/ (oracle/fetch_ref.py)
and the batched composition oracle of SURVEY.md section 8(c) built from its pieces:

    clamp -> integrators.rollout_trajectory_differentiable (src/integrators.py:192-258)
          -> quadratic horizon cost (src/mpc_controller.py:75-114) -> .backward() -> torch.optim.Adam

which equals B independent runs of MPCController.compute_control (src/mpc_controller.py:143-209) /
MPCControllerCanonical.optimize_control (src/mpc_controller_canonical.py:163-228) because instances
are independent and Adam is elementwise.  Used by `bench.py --impl reference`, bench.py's
`cpu_baseline` leg and tests/; never imported by phnn_mpc_b200/.  Reads only oracle/_ref (never
/root/reference), so it also runs on the GPU box.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_mods = None


def available():
    from . import fetch_ref
    return fetch_ref.available()


def modules():
    """import the staged reference modules (pHNN, pHNN_canonical, integrators, controllers)"""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/fetch_ref.py` where /root/reference exists")
    sys.dont_write_bytecode = True
    src = os.path.join(REF, "src")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("pHNN", "pHNN_canonical", "NN", "mass_matrix", "coordinate_transforms", "integrators",
                      "mpc_controller", "mpc_controller_canonical", "cartpole_simulator")}
    sys.path.insert(0, src)
    try:
        import pHNN as m_phnn
        import pHNN_canonical as m_canon
        import integrators as m_int
        import mpc_controller as m_mpc
        import mpc_controller_canonical as m_mpcc
        import cartpole_simulator as m_sim
        for m in (m_phnn, m_canon, m_int, m_mpc, m_mpcc, m_sim):
            assert os.path.abspath(m.__file__).startswith(os.path.abspath(src)), m.__file__
        _mods = dict(pHNN=m_phnn.pHNN, pHNN_Canonical=m_canon.pHNN_Canonical, integrators=m_int,
                     MPCController=m_mpc.MPCController, mpc_controller_canonical=m_mpcc,
                     CartPoleSimulator=m_sim.CartPoleSimulator)
    finally:
        sys.path.remove(src)
        # leave the reference modules registered under private names only: tests may import the drop-ins next
        for k in ("pHNN", "pHNN_canonical", "NN", "mass_matrix", "coordinate_transforms", "integrators",
                  "mpc_controller", "mpc_controller_canonical", "cartpole_simulator"):
            mod = sys.modules.pop(k, None)
            if mod is not None:
                sys.modules["_phnn_ref_" + k] = mod
        sys.modules.update(saved)
    return _mods


def config_path(name):
    return os.path.join(REF, {"cartpole": "cartpole_mpc_config.yaml", "pole": "pole_stabilization_config.yaml",
                              "pendulum": "pendulum_config.yaml"}[name])


def load_model(sd, kind):
    """reference nn.Module of `kind` ('phnn' | 'canonical') with the state_dict `sd` (numpy arrays);
    the hidden width and state dimension are read from the weights, the rest from the reference's own YAML."""
    import torch
    import yaml
    M = modules()
    n = int(sd["H_net.net.0.weight"].shape[1])
    h = int(sd["H_net.net.0.weight"].shape[0])
    cfg = yaml.safe_load(open(config_path("pendulum" if n == 2 else "cartpole")))
    cfg["model"]["H_mlp"]["hidden_sizes"] = [h, h]
    if "R_mlp" in cfg["model"]:
        cfg["model"]["R_mlp"]["hidden_sizes"] = [h]
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    try:
        model = (M["pHNN_Canonical"] if kind == "canonical" else M["pHNN"])(f.name)
    finally:
        os.unlink(f.name)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)).clone() for k, v in sd.items()})
    model.eval()
    return model


def composition_solve(model, x0, U0, dt, integrator, Q, R, xt, umin, umax, lr, iters, return_mode="last",
                      want_grads=False):
    """B-instance MPC solve composed of reference pieces (see module docstring).
    Returns dict(U [B,H,1] per return_mode, hist [iters,B], best [B], grad0, grad_last)."""
    import torch
    M = modules()
    rollout = M["integrators"].rollout_trajectory_differentiable
    x0 = torch.as_tensor(x0, dtype=torch.float32)
    U0 = torch.as_tensor(U0, dtype=torch.float32)
    Q = torch.as_tensor(Q, dtype=torch.float32)
    R = torch.as_tensor(R, dtype=torch.float32).reshape(1, 1)
    xt = torch.as_tensor(xt, dtype=torch.float32)
    U = U0.clone().requires_grad_(True)
    opt = torch.optim.Adam([U], lr=lr)
    hist, g0, gl = [], None, None
    best = torch.full((x0.shape[0],), float("inf"))
    Ubest = torch.clamp(U0, umin, umax).clone()
    for it in range(iters):
        opt.zero_grad()
        Uc = torch.clamp(U, umin, umax)
        traj = rollout(model, x0.clone().requires_grad_(True), Uc, dt, integrator)
        e = traj - xt
        cost = torch.einsum("bti,ij,btj->b", e, Q, e) + torch.einsum("bti,ij,btj->b", Uc, R, Uc)
        cost.sum().backward()
        if want_grads:
            if it == 0:
                g0 = U.grad.detach().clone()
            gl = U.grad.detach().clone()
        c = cost.detach()
        hist.append(c.clone())
        better = c < best
        best = torch.where(better, c, best)
        Ubest[better] = Uc.detach()[better]
        opt.step()
    Ulast = torch.clamp(U.detach(), umin, umax)
    out = dict(U=(Ulast if return_mode == "last" else Ubest).numpy(), U_last=Ulast.numpy(), U_best=Ubest.numpy(),
               best=best.numpy(), hist=torch.stack(hist).numpy() if hist else np.zeros((0, x0.shape[0]), np.float32))
    if want_grads:
        out.update(grad0=g0.numpy(), grad_last=gl.numpy())
    return out


def time_solve(model, x0, H, dt, integrator, cost, lr, iters, threads, return_mode="last"):
    """wall-clock seconds of ONE full composition solve of x0.shape[0] instances at `threads` torch threads"""
    import torch
    torch.set_num_threads(int(threads))
    Q = np.diag(np.asarray(cost["Q"], np.float32))
    U0 = np.zeros((x0.shape[0], H, 1), np.float32)
    t0 = time.perf_counter()
    composition_solve(model, x0, U0, dt, integrator, Q, cost["R"][0], np.zeros(4, np.float32), cost["u_min"],
                      cost["u_max"], lr, iters, return_mode)
    return time.perf_counter() - t0
