"""
tests/golden/make_golden.py -- regenerates the golden vectors under tests/golden/.

Runs ONLY in the build container: it imports the unmodified reference modules from
/root/reference/src (read-only) and records, with fixed seeds, what the reference's own
PyTorch-autograd implementation returns on the hot path.  The GPU box never runs this; it reads
the committed .npz files.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Every fixture stores the model's state_dict (prefix "sd/") so no checkpoint has to travel.
Reference entry points exercised (paths under /root/reference):
  src/pHNN.py:52-100, src/pHNN_canonical.py:172-273, src/integrators.py:13-258,
  src/mpc_controller.py:75-209, src/mpc_controller_canonical.py:91-273.
"""
import os
import sys
import tempfile

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True

import numpy as np
import torch
import yaml

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CFG = os.path.join(REPO, "configs")

from pHNN import pHNN  # noqa: E402  (reference)
from pHNN_canonical import pHNN_Canonical  # noqa: E402
from integrators import rollout_trajectory, rollout_trajectory_differentiable  # noqa: E402
from mpc_controller import MPCController  # noqa: E402
from mpc_controller_canonical import create_mpc_controller  # noqa: E402

torch.set_num_threads(1)


def sd_np(model):
    return {"sd/" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def wide_cfg(hidden):
    cfg = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    cfg["model"]["H_mlp"]["hidden_sizes"] = [hidden, hidden]
    cfg["model"]["R_mlp"]["hidden_sizes"] = [hidden]
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    return f.name


def fwd_and_vjp(model, x, u, v):
    """model forward, and (d dx/d x)^T v, (d dx/d u)^T v by autograd."""
    x = x.clone().requires_grad_(True)
    u = u.clone().requires_grad_(True)
    out = model(x, u)
    dx, H = out[0], out[1]
    gx, gu = torch.autograd.grad((dx * v).sum(), [x, u])
    return dx.detach(), H.detach(), gx, gu


def batched_cost(traj, Uc, Q, R, xt):
    e = traj - xt
    sc = torch.einsum("bti,ij,btj->b", e, Q, e)
    cc = torch.einsum("bti,ij,btj->b", Uc, R, Uc)
    return sc + cc


def composition_solve(model, x0, U0, dt, integrator, Q, R, xt, umin, umax, lr, iters, mode):
    """Batched composition of reference pieces: clamp ->
    rollout_trajectory_differentiable -> quadratic cost -> backward -> torch.optim.Adam."""
    U = U0.clone().requires_grad_(True)
    opt = torch.optim.Adam([U], lr=lr)
    hist, grads = [], []
    best = torch.full((x0.shape[0],), float("inf"))
    Ubest = torch.clamp(U0, umin, umax).clone()
    for _ in range(iters):
        opt.zero_grad()
        Uc = torch.clamp(U, umin, umax)
        traj = rollout_trajectory_differentiable(model, x0.clone().requires_grad_(True), Uc, dt, integrator)
        cost = batched_cost(traj, Uc, Q, R, xt)
        cost.sum().backward()
        grads.append(U.grad.detach().clone())
        hist.append(cost.detach().clone())
        better = cost.detach() < best
        best = torch.where(better, cost.detach(), best)
        Ubest[better] = Uc.detach()[better]
        opt.step()
    Ulast = torch.clamp(U.detach(), umin, umax)
    return dict(U_last=Ulast.numpy(), U_best=Ubest.numpy(), best=best.numpy(), hist=torch.stack(hist).numpy(),
                grad0=grads[0].numpy(), grad_last=grads[-1].numpy())


def gen_pendulum():
    model = pHNN(os.path.join(CFG, "pendulum_phnn.yaml"))
    model.load_state_dict(torch.load(os.path.join(REF, "pendulum_pHNN_weights.pth"), map_location="cpu"))
    model.eval()
    out = sd_np(model)
    # Appendix C anchors
    x = torch.tensor([[1.0, 0.5], [-2.0, 0.25]])
    u = torch.tensor([[0.3], [-1.0]])
    v = torch.tensor([[0.7, -0.2], [0.1, 1.3]])
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(anchor_x=x.numpy(), anchor_u=u.numpy(), anchor_v=v.numpy(), anchor_dx=dx.numpy(), anchor_H=H.numpy(),
               anchor_gx=gx.numpy(), anchor_gu=gu.numpy())
    U10 = u[:, None, :].repeat(1, 10, 1)
    for integ in ("rk4", "euler"):
        tr, en = rollout_trajectory_differentiable(model, x.clone().requires_grad_(True), U10, 0.05, integ, True)
        out["anchor_traj_" + integ] = tr.detach().numpy()
        out["anchor_en_" + integ] = en.detach().numpy()
        tr2, en2 = rollout_trajectory(model, x.clone().requires_grad_(True), U10, 0.05, integ)
        out["anchor_traj2_" + integ] = tr2.detach().numpy()
        out["anchor_en2_" + integ] = en2.detach().numpy()
    Ug = U10.clone().requires_grad_(True)
    tr = rollout_trajectory_differentiable(model, x.clone().requires_grad_(True), Ug, 0.05, "rk4")
    Jv = (tr ** 2).sum()
    Jv.backward()
    out.update(anchor_J=np.float32(Jv.item()), anchor_dJdU=Ug.grad.numpy())
    # config-2 style inputs, small
    g = torch.Generator().manual_seed(1)
    B, T = 64, 100
    th = (torch.rand(B, generator=g) * 2 - 1) * np.pi
    om = torch.rand(B, generator=g) * 2 - 1
    x0 = torch.stack([th, om], 1)
    U = (torch.rand(B, T, 1, generator=g) * 4 - 2)
    for integ in ("rk4", "euler"):
        tr, en = rollout_trajectory_differentiable(model, x0.clone().requires_grad_(True), U, 0.05, integ, True)
        out["cfg2_traj_" + integ] = tr.detach().numpy()
        out["cfg2_en_" + integ] = en.detach().numpy()
    out.update(cfg2_x0=x0.numpy(), cfg2_U=U.numpy())
    # random forward / vjp points
    xr = torch.randn(32, 2, generator=g) * 2
    ur = torch.randn(32, 1, generator=g)
    vr = torch.randn(32, 2, generator=g)
    dx, H, gx, gu = fwd_and_vjp(model, xr, ur, vr)
    out.update(rand_x=xr.numpy(), rand_u=ur.numpy(), rand_v=vr.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy())
    # quadratic-cost gradient through the learned-G model (euler + rk4)
    Q = torch.diag(torch.tensor([5.0, 1.0]))
    R = torch.tensor([[0.1]])
    xt = torch.tensor([0.5, 0.0])
    res = {}
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0[:8], U[:8, :15] * 1.2, 0.05, integ, Q, R, xt, -2.0, 2.0, 0.05, 6, "last")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
    out.update(mpc_x0=x0[:8].numpy(), mpc_U0=(U[:8, :15] * 1.2).numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(),
               mpc_xt=xt.numpy(), mpc_bounds=np.array([-2.0, 2.0], np.float32), mpc_lr=np.float32(0.05),
               mpc_dt=np.float32(0.05))
    np.savez(os.path.join(HERE, "pendulum.npz"), **out)
    print("pendulum: anchor dx", dx[:1].numpy() if False else out["anchor_dx"], "J", out["anchor_J"])


def cartpole_points(g, B):
    x = (torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])
    u = (torch.rand(B, 1, generator=g) * 2 - 1) * 10
    v = torch.randn(B, 4, generator=g)
    return x, u, v


def gen_cartpole(hidden, seed, name):
    cfg_path = os.path.join(CFG, "cartpole_phnn.yaml") if hidden == 128 else wide_cfg(hidden)
    torch.manual_seed(seed)
    model = pHNN(cfg_path)
    model.eval()
    cfg = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    mpc = cfg["mpc"]
    out = sd_np(model)
    g = torch.Generator().manual_seed(7)
    x, u, v = cartpole_points(g, 32)
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(rand_x=x.numpy(), rand_u=u.numpy(), rand_v=v.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy())
    # wide-range states as in data/cartpole_training_data.pt
    xw = torch.randn(16, 4, generator=g) * torch.tensor([5.0, 3.0, 4.0, 4.0])
    dxw, Hw, gxw, guw = fwd_and_vjp(model, xw, u[:16], v[:16])
    out.update(wide_x=xw.numpy(), wide_dx=dxw.numpy(), wide_H=Hw.numpy(), wide_gx=gxw.numpy())
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.tensor([[mpc["R_diag"][0]]])
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    B, Hh, iters = 8, 12, 5
    x0 = x[:B]
    U0 = (torch.rand(B, Hh, 1, generator=g) * 2 - 1) * 18.0       # some entries beyond the +-15 bounds
    out.update(mpc_x0=x0.numpy(), mpc_U0=U0.numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(), mpc_xt=xt.numpy(),
               mpc_bounds=np.array([mpc["u_min"], mpc["u_max"]], np.float32), mpc_lr=np.float32(mpc["learning_rate"]),
               mpc_dt=np.float32(dt))
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0, U0, dt, integ, Q, R, xt, mpc["u_min"], mpc["u_max"], mpc["learning_rate"],
                                iters, "last")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
        tr = rollout_trajectory_differentiable(model, x0.clone().requires_grad_(True),
                                               torch.clamp(U0, mpc["u_min"], mpc["u_max"]), dt, integ)
        out["mpc_%s_traj0" % integ] = tr.detach().numpy()
    if hidden == 128:
        # the real controller, single instance, YAML parameters (config 1)
        ctrl = MPCController(model, mpc["horizon"], dt, mpc["Q_diag"], mpc["R_diag"][0], mpc["x_target"], mpc["u_min"],
                             mpc["u_max"], optimizer_type="Adam", lr=mpc["learning_rate"],
                             max_iterations=mpc["optimizer_steps"])
        xs = np.array([[0.0, 0.1, 0.0, 0.0], [0.3, -0.12, 0.2, -0.4], [-0.5, 0.2, -0.3, 0.6]], np.float64)
        us = [ctrl.compute_control(s) for s in xs]
        out.update(ctrl_x=xs, ctrl_u=np.stack(us))
        # with soft state bounds (the optional barrier term, src/mpc_controller.py:96-107)
        ctrl_b = MPCController(model, 8, dt, mpc["Q_diag"], mpc["R_diag"][0], mpc["x_target"], mpc["u_min"], mpc["u_max"],
                               x_min=[-0.2, -0.05, -0.1, -0.2], x_max=[0.2, 0.05, 0.1, 0.2], optimizer_type="Adam",
                               lr=mpc["learning_rate"], max_iterations=6)
        out.update(ctrlb_u=np.stack([ctrl_b.compute_control(s) for s in xs]))
        Ub = torch.zeros(8, 1)
        st = ctrl_b.rollout_dynamics(torch.tensor(xs[1], dtype=torch.float32), Ub)
        out.update(ctrlb_cost0=np.float32(ctrl_b.compute_cost(st, Ub).item()))
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, "done; H[0..2] =", out["rand_H"][:3])


def gen_cartpole_dropout():
    """Cart-pole pHNN whose MLPs were built with dropout = 0.1 (src/NN.py:16-25: nn.Dropout after every activation, so the
    Linear layers sit at net.0 / net.3 / net.6), in EVAL mode -- the mode MPCController puts the model in
    (src/mpc_controller.py:44): forward, autograd VJP and the composition oracle, plus the real controller on one state."""
    cfgd = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    cfgd["model"]["H_mlp"]["dropout"] = 0.1
    cfgd["model"]["R_mlp"]["dropout"] = 0.1
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfgd, f)
    f.close()
    torch.manual_seed(21)
    model = pHNN(f.name)
    os.unlink(f.name)
    model.eval()
    mpc = cfgd["mpc"]
    out = sd_np(model)
    g = torch.Generator().manual_seed(9)
    x, u, v = cartpole_points(g, 32)
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(rand_x=x.numpy(), rand_u=u.numpy(), rand_v=v.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy())
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.tensor([[mpc["R_diag"][0]]])
    xt = torch.tensor(mpc["x_target"])
    dt = cfgd["cartpole"]["dt"]
    B, Hh, iters = 8, 12, 5
    x0 = x[:B]
    U0 = (torch.rand(B, Hh, 1, generator=g) * 2 - 1) * 18.0
    out.update(mpc_x0=x0.numpy(), mpc_U0=U0.numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(), mpc_xt=xt.numpy(),
               mpc_bounds=np.array([mpc["u_min"], mpc["u_max"]], np.float32), mpc_lr=np.float32(mpc["learning_rate"]),
               mpc_dt=np.float32(dt))
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0, U0, dt, integ, Q, R, xt, mpc["u_min"], mpc["u_max"], mpc["learning_rate"],
                                iters, "last")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
    ctrl = MPCController(model, mpc["horizon"], dt, mpc["Q_diag"], mpc["R_diag"][0], mpc["x_target"], mpc["u_min"],
                         mpc["u_max"], optimizer_type="Adam", lr=mpc["learning_rate"], max_iterations=mpc["optimizer_steps"])
    xs = np.array([[0.0, 0.1, 0.0, 0.0]], np.float64)
    out.update(ctrl_x=xs, ctrl_u=np.stack([ctrl.compute_control(s) for s in xs]))
    np.savez(os.path.join(HERE, "cartpole_h128_dropout.npz"), **out)
    print("cartpole_h128_dropout done; keys:", sorted(k for k in out if k.startswith("sd/H_net")))


def gen_canonical_nobias():
    """Canonical pHNN whose H_mlp is configured with ``bias: false`` (src/NN.py:19,25: Linear layers without bias), by the
    reference's constructor: forward, autograd VJP and the composition oracle."""
    cfgd = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    cfgd["model"]["H_mlp"]["bias"] = False
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfgd, f)
    f.close()
    torch.manual_seed(33)
    model = pHNN_Canonical(f.name)
    os.unlink(f.name)
    model.eval()
    with torch.no_grad():
        model.M_net.log_a.copy_(torch.tensor(0.1))
        model.M_net.b.copy_(torch.tensor(0.25))
        model.R_diag_raw.copy_(torch.tensor([0.2, -0.1, 0.4, 0.9]))
    cfg = yaml.safe_load(open(os.path.join(CFG, "pole_stabilization.yaml")))
    mpc = cfg["mpc"]
    out = sd_np(model)
    g = torch.Generator().manual_seed(13)
    x, u, v = cartpole_points(g, 32)
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(rand_x=x.numpy(), rand_u=u.numpy(), rand_v=v.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy())
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.diag(torch.tensor(mpc["R_diag"]))
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    B, Hh, iters = 8, 10, 6
    x0 = torch.tensor([0.0, 0.05, 0.0, 0.0]) + (torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])
    U0 = (torch.rand(B, Hh, 1, generator=g) * 2 - 1) * 35.0
    out.update(mpc_x0=x0.numpy(), mpc_U0=U0.numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(), mpc_xt=xt.numpy(),
               mpc_bounds=np.array([mpc["u_min"], mpc["u_max"]], np.float32), mpc_lr=np.float32(mpc["learning_rate"]),
               mpc_dt=np.float32(dt))
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0, U0, dt, integ, Q, R, xt, mpc["u_min"], mpc["u_max"], mpc["learning_rate"],
                                iters, "best")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
    np.savez(os.path.join(HERE, "canonical_nobias.npz"), **out)
    print("canonical_nobias done; keys:", sorted(k for k in out if k.startswith("sd/H_net")))


def gen_lbfgs():
    """The LBFGS branch of MPCController.compute_control (src/mpc_controller.py:169-170: torch.optim.LBFGS(lr, max_iter=20),
    stepped max_iterations times) on the cartpole_h128 model: three states, a short configuration (H=10, 3 outer steps)."""
    z = np.load(os.path.join(HERE, "cartpole_h128.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model = pHNN(os.path.join(CFG, "cartpole_phnn.yaml"))
    model.load_state_dict(sd)
    model.eval()
    mpc = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))["mpc"]
    xs = np.array([[0.0, 0.1, 0.0, 0.0], [0.3, -0.12, 0.2, -0.4], [-0.5, 0.2, -0.3, 0.6]], np.float64)
    out = {"x": xs}
    for lr, tag in ((0.05, "a"), (0.2, "b")):
        ctrl = MPCController(model, 10, 0.02, mpc["Q_diag"], mpc["R_diag"][0], mpc["x_target"], mpc["u_min"], mpc["u_max"],
                             optimizer_type="LBFGS", lr=lr, max_iterations=3)
        out["u_" + tag] = np.stack([ctrl.compute_control(s) for s in xs])
        out["lr_" + tag] = np.float32(lr)
        # the cost the branch reaches: the same loop again with the controller's own rollout_dynamics / compute_cost
        # (compute_control does not return the sequence); bit-identical to the call above (deterministic CPU arithmetic)
        Js = []
        for s_ in xs:
            x0 = torch.tensor(s_, dtype=torch.float32)
            cs = torch.zeros(10, 1, requires_grad=True)
            opt = torch.optim.LBFGS([cs], lr=lr, max_iter=20)

            def closure():
                opt.zero_grad()
                c = torch.clamp(cs, mpc["u_min"], mpc["u_max"])
                J = ctrl.compute_cost(ctrl.rollout_dynamics(x0, c), c)
                J.backward()
                return J
            for _ in range(3):
                opt.step(closure)
            c = torch.clamp(cs.detach(), mpc["u_min"], mpc["u_max"])   # (the model needs autograd for grad H: no no_grad here)
            assert abs(float(c[0]) - float(out["u_" + tag][len(Js)][0])) < 1e-6
            Js.append(float(ctrl.compute_cost(ctrl.rollout_dynamics(x0, c), c).item()))
        out["J_" + tag] = np.array(Js, np.float32)
    np.savez(os.path.join(HERE, "lbfgs.npz"), **out)
    print("lbfgs done; u =", out["u_a"].ravel(), out["u_b"].ravel())


def gen_canonical():
    torch.manual_seed(0)
    model = pHNN_Canonical(os.path.join(CFG, "cartpole_phnn.yaml"))
    model.eval()
    # perturb the physical parameters so every derivative path is exercised
    with torch.no_grad():
        model.M_net.log_a.copy_(torch.tensor(0.25))
        model.M_net.b.copy_(torch.tensor(0.35))
        model.M_net.log_c.copy_(torch.tensor(-0.4))
        model.R_diag_raw.copy_(torch.tensor([0.1, -0.3, 0.5, 1.2]))
    cfg = yaml.safe_load(open(os.path.join(CFG, "pole_stabilization.yaml")))
    mpc = cfg["mpc"]
    out = sd_np(model)
    g = torch.Generator().manual_seed(3)
    x, u, v = cartpole_points(g, 32)
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(rand_x=x.numpy(), rand_u=u.numpy(), rand_v=v.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy())
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.diag(torch.tensor(mpc["R_diag"]))
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    B, Hh, iters = 8, 10, 6
    x0 = torch.tensor([0.0, 0.05, 0.0, 0.0]) + (torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])
    U0 = (torch.rand(B, Hh, 1, generator=g) * 2 - 1) * 35.0
    out.update(mpc_x0=x0.numpy(), mpc_U0=U0.numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(), mpc_xt=xt.numpy(),
               mpc_bounds=np.array([mpc["u_min"], mpc["u_max"]], np.float32), mpc_lr=np.float32(mpc["learning_rate"]),
               mpc_dt=np.float32(dt))
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0, U0, dt, integ, Q, R, xt, mpc["u_min"], mpc["u_max"], mpc["learning_rate"],
                                iters, "best")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
    # the real canonical controller (config 3, B=1): cold start then a warm start
    ctrl = create_mpc_controller(model, cfg)
    xs = x0[:3].numpy().astype(np.float64)
    us, seqs, costs = [], [], []
    for s in xs:
        u1, info1 = ctrl.control(s)
        u2, info2 = ctrl.control(s + 0.01, info1["u_sequence"])
        us.append(np.stack([u1, u2]))
        seqs.append(np.stack([info1["u_sequence"], info2["u_sequence"]]))
        costs.append(np.stack([np.array(info1["optimization"]["costs"], np.float32),
                               np.array(info2["optimization"]["costs"], np.float32)]))
    out.update(ctrl_x=xs, ctrl_u=np.stack(us), ctrl_seq=np.stack(seqs), ctrl_costs=np.stack(costs))
    np.savez(os.path.join(HERE, "canonical.npz"), **out)
    print("canonical done; dx[0] =", out["rand_dx"][0])




def gen_canonical_constM():
    """Canonical pHNN with the CONSTANT mass matrix of MassMatrixNetwork (src/mass_matrix.py:15-216, mass_type
    'constant': M = L L^T, exact inverse, no dependence on q), built by the reference's own constructor branch
    (src/pHNN_canonical.py:79-86) from the cart-pole YAML with ``mass_matrix.type: constant``: forward, autograd VJP,
    cost + dJ/dU and Adam trajectories of the composition oracle."""
    import tempfile
    cfgd = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    cfgd["model"]["mass_matrix"] = {"type": "constant", "init_scale": 1.0}
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        yaml.safe_dump(cfgd, f)
        path = f.name
    torch.manual_seed(0)
    model = pHNN_Canonical(path)
    os.unlink(path)
    model.eval()
    with torch.no_grad():   # a non-trivial, non-diagonal factor (the upper triangle is ignored by the reference: tril)
        model.M_net.L_tril.copy_(torch.tensor([[0.9, 0.7], [0.35, 0.4]]))
        model.R_diag_raw.copy_(torch.tensor([0.1, -0.3, 0.5, 1.2]))
    cfg = yaml.safe_load(open(os.path.join(CFG, "pole_stabilization.yaml")))
    mpc = cfg["mpc"]
    out = sd_np(model)
    g = torch.Generator().manual_seed(5)
    x, u, v = cartpole_points(g, 32)
    dx, H, gx, gu = fwd_and_vjp(model, x, u, v)
    out.update(rand_x=x.numpy(), rand_u=u.numpy(), rand_v=v.numpy(), rand_dx=dx.numpy(), rand_H=H.numpy(),
               rand_gx=gx.numpy(), rand_gu=gu.numpy(), M=model.M_net(x[:1, :2])[0].detach().numpy())
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.diag(torch.tensor(mpc["R_diag"]))
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    B, Hh, iters = 8, 10, 6
    x0 = torch.tensor([0.0, 0.05, 0.0, 0.0]) + (torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])
    U0 = (torch.rand(B, Hh, 1, generator=g) * 2 - 1) * 35.0
    out.update(mpc_x0=x0.numpy(), mpc_U0=U0.numpy(), mpc_Q=Q.numpy(), mpc_R=R.numpy(), mpc_xt=xt.numpy(),
               mpc_bounds=np.array([mpc["u_min"], mpc["u_max"]], np.float32), mpc_lr=np.float32(mpc["learning_rate"]),
               mpc_dt=np.float32(dt))
    for integ in ("euler", "rk4"):
        res = composition_solve(model, x0, U0, dt, integ, Q, R, xt, mpc["u_min"], mpc["u_max"], mpc["learning_rate"],
                                iters, "best")
        for k, val in res.items():
            out["mpc_%s_%s" % (integ, k)] = val
    np.savez(os.path.join(HERE, "canonical_constM.npz"), **out)
    print("canonical_constM done; M =", out["M"].tolist(), "dx[0] =", out["rand_dx"][0])


def gen_closed_loop():
    """Closed loop of the reference pieces: CartPoleSimulator <-> controller, the
    loop body of scripts/run_cartpole_mpc.py:121-176 and scripts/run_mpc_canonical.py:55-95 (warm start)."""
    from cartpole_simulator import CartPoleSimulator  # reference
    steps = 6
    out = {}
    # --- pHNN + MPCController, cold start every step (config 1) ---
    torch.manual_seed(0)
    model = pHNN(os.path.join(CFG, "cartpole_phnn.yaml"))
    model.eval()
    cfg = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    mpc = cfg["mpc"]
    ctrl = MPCController(model, mpc["horizon"], cfg["cartpole"]["dt"], mpc["Q_diag"], mpc["R_diag"][0], mpc["x_target"],
                         mpc["u_min"], mpc["u_max"], optimizer_type="Adam", lr=mpc["learning_rate"],
                         max_iterations=mpc["optimizer_steps"])
    x0s = np.array([[0.0, 0.1, 0.0, 0.0], [0.3, -0.12, 0.2, -0.4], [0.05, 0.02, -0.04, 0.03]])
    tol = np.array(cfg["stability"]["tolerance"])
    traj, ctr, hs, stable = [], [], [], []
    for x0 in x0s:
        sim = CartPoleSimulator(cfg["cartpole"]["dt"])
        sim.reset(x0)
        state = x0.copy()
        states, controls, hams, within = [state.copy()], [], [], []
        for _ in range(steps):
            u = ctrl.compute_control(state)
            controls.append(float(u[0]))
            st = torch.tensor(state, dtype=torch.float32, requires_grad=True).unsqueeze(0)
            _, H = ctrl.model(st, torch.tensor([[float(u[0])]], dtype=torch.float32))
            hams.append(H.detach().item())
            within.append(bool(np.all(np.abs(state - np.array(mpc["x_target"])) <= tol)))
            state, done = sim.step(u)
            states.append(state.copy())
        traj.append(np.array(states)); ctr.append(np.array(controls)); hs.append(np.array(hams)); stable.append(np.array(within))
    out.update({"sd/" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()})
    out.update(cl_x0=x0s, cl_traj=np.stack(traj), cl_u=np.stack(ctr).astype(np.float32), cl_H=np.stack(hs).astype(np.float32),
               cl_within=np.stack(stable))
    # plant alone: 50 steps under a fixed control pattern, float64
    sim = CartPoleSimulator(0.02)
    sim.reset(np.array([0.1, 0.2, -0.3, 0.4]))
    ps = [sim.get_state()]
    us = (np.sin(np.arange(50) * 0.7) * 12.0).astype(np.float32)
    dones = []
    for u in us:
        s, d = sim.step(np.array([u], np.float32))
        ps.append(s); dones.append(d)
    out.update(plant_traj=np.array(ps), plant_u=us, plant_done=np.array(dones))
    np.savez(os.path.join(HERE, "closed_loop.npz"), **out)
    # --- canonical + MPCControllerCanonical, warm start from the shifted previous plan ---
    torch.manual_seed(0)
    cm = pHNN_Canonical(os.path.join(CFG, "cartpole_phnn.yaml"))
    cm.eval()
    pcfg = yaml.safe_load(open(os.path.join(CFG, "pole_stabilization.yaml")))
    cc = create_mpc_controller(cm, pcfg)
    x0c = np.array([[0.0, 0.05, 0.0, 0.0], [0.2, -0.08, 0.1, 0.2]])
    ctraj, cctr = [], []
    for x0 in x0c:
        sim = CartPoleSimulator(pcfg["cartpole"]["dt"])
        sim.reset(x0)
        state, u_prev = x0.copy(), None
        states, controls = [state.copy()], []
        for _ in range(steps):
            u, info = cc.control(state, u_prev)
            u_prev = info["u_sequence"]
            controls.append(float(u[0]))
            state, done = sim.step(u)
            states.append(state.copy())
        ctraj.append(np.array(states)); cctr.append(np.array(controls))
    outc = {"sd/" + k: v.detach().numpy().copy() for k, v in cm.state_dict().items()}
    outc.update(cl_x0=x0c, cl_traj=np.stack(ctraj), cl_u=np.stack(cctr).astype(np.float32))
    np.savez(os.path.join(HERE, "closed_loop_canonical.npz"), **outc)
    print("closed loop done; u[0] =", out["cl_u"][0], "canon u[0] =", outc["cl_u"][0])


def gen_cfg4_shape():
    """BASELINE cfg4 at a reference-runnable size: the cfg4 model (cartpole_h256.npz weights), the first 64 of the
    benchmark's own instances (bench.make_inputs(., 'phnn', 7)), H=50, RK4, 20 Adam iterations, cold start --
    exactly the job bench.py times per instance (src/integrators.py:192-258 + src/mpc_controller.py:143-209 composed
    as in SURVEY.md section 8c).  About two minutes of reference CPU time."""
    z = np.load(os.path.join(HERE, "cartpole_h256.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model = pHNN(wide_cfg(256))
    model.load_state_dict(sd)
    model.eval()
    cfg = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    mpc = cfg["mpc"]
    g = torch.Generator().manual_seed(7)
    B, Hh, iters = 64, 50, 20
    x0 = ((torch.rand(65536, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).float()[:B]
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.tensor([[mpc["R_diag"][0]]])
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    res = composition_solve(model, x0, torch.zeros(B, Hh, 1), dt, "rk4", Q, R, xt, mpc["u_min"], mpc["u_max"],
                            mpc["learning_rate"], iters, "last")
    out = {"x0": x0.numpy(), "H": np.int32(Hh), "iters": np.int32(iters), "dt": np.float32(dt),
           "lr": np.float32(mpc["learning_rate"]), "Q": Q.numpy(), "R": R.numpy(), "xt": xt.numpy(),
           "bounds": np.array([mpc["u_min"], mpc["u_max"]], np.float32)}
    for k, val in res.items():
        out["rk4_" + k] = val
    np.savez(os.path.join(HERE, "cfg4_shape.npz"), **out)
    print("cfg4_shape done; cost[0] first/last =", res["hist"][0, 0], res["hist"][-1, 0])


def gen_cfg5_shape():
    """BASELINE cfg5 horizons at a reference-runnable size: the cfg4/cfg5 model (cartpole_h256.npz weights), the first 16
    of the benchmark's own instances, H = 100 and H = 200, RK4, 4 Adam iterations, cold start (the long horizons were
    pinned only through the oracle before).  About a minute of reference CPU time."""
    z = np.load(os.path.join(HERE, "cartpole_h256.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model = pHNN(wide_cfg(256))
    model.load_state_dict(sd)
    model.eval()
    cfg = yaml.safe_load(open(os.path.join(CFG, "cartpole_phnn.yaml")))
    mpc = cfg["mpc"]
    g = torch.Generator().manual_seed(7)
    B, iters = 16, 4
    x0 = ((torch.rand(65536, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])).float()[:B]
    Q = torch.diag(torch.tensor(mpc["Q_diag"]))
    R = torch.tensor([[mpc["R_diag"][0]]])
    xt = torch.tensor(mpc["x_target"])
    dt = cfg["cartpole"]["dt"]
    out = {"x0": x0.numpy(), "iters": np.int32(iters), "dt": np.float32(dt), "lr": np.float32(mpc["learning_rate"]),
           "Q": Q.numpy(), "R": R.numpy(), "xt": xt.numpy(), "bounds": np.array([mpc["u_min"], mpc["u_max"]], np.float32)}
    for Hh in (100, 200):
        res = composition_solve(model, x0, torch.zeros(B, Hh, 1), dt, "rk4", Q, R, xt, mpc["u_min"], mpc["u_max"],
                                mpc["learning_rate"], iters, "last")
        for k, val in res.items():
            out["h%d_%s" % (Hh, k)] = val
        print("cfg5_shape H=%d done; cost[0] first/last =" % Hh, res["hist"][0, 0], res["hist"][-1, 0])
    np.savez(os.path.join(HERE, "cfg5_shape.npz"), **out)


def gen_train():
    """Weight gradients of the reference's own training losses, by autograd:
    scripts/train_cartpole_phnn.py:108-178 (pHNN: Euler unroll, position MSE + angle cosine + velocity MSE; the energy
    anchor term is left out here because it does not involve the rollout) and scripts/train_cartpole_phnn_canonical.py:
    83-196 (canonical: Euler unroll, position + velocity-reconstruction loss), plus an RK4 trajectory-matching loss
    through integrators.rollout_trajectory_differentiable.  Batch: 16 windows of 16 steps of data/cartpole_training_data.pt."""
    from coordinate_transforms import split_state
    data = torch.load(os.path.join(REF, "data", "cartpole_training_data.pt"))
    X, Uall = data["states"] if isinstance(data, dict) and "states" in data else None, None
    if X is None:
        keys = list(data.keys()) if isinstance(data, dict) else []
        raise RuntimeError("unexpected training data layout: %s" % keys)
    Uall = data["controls"] if "controls" in data else data["actions"]
    g = torch.Generator().manual_seed(11)
    idx = torch.randint(0, X.shape[0], (16,), generator=g)
    t0 = torch.randint(0, X.shape[1] - 16, (16,), generator=g)
    xb = torch.stack([X[i, s:s + 16] for i, s in zip(idx.tolist(), t0.tolist())]).float()     # [16,16,4]
    ub = torch.stack([Uall[i, s:s + 16] for i, s in zip(idx.tolist(), t0.tolist())]).float()  # [16,16,1]
    dt = 0.02
    out = {"x_batch": xb.numpy(), "u_batch": ub.numpy(), "dt": np.float32(dt)}

    def grads_of(model, loss, prefix):
        model.zero_grad()
        loss.backward()
        for k, p_ in model.named_parameters():
            out[prefix + "/" + k] = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).detach().numpy().copy()
        out[prefix + "/loss"] = np.float32(loss.item())

    # --- pHNN (scripts/train_cartpole_phnn.py:108-160 without the energy anchor) ---
    torch.manual_seed(0)
    model = pHNN(os.path.join(CFG, "cartpole_phnn.yaml"))
    loss_fn = torch.nn.MSELoss()
    x0b = xb[:, 0, :].clone().requires_grad_(True)
    Xp = [x0b]
    for t in range(xb.shape[1] - 1):
        dx, _ = model(Xp[-1], ub[:, t, :])
        Xp.append(Xp[-1] + dt * dx)
    Xp = torch.stack(Xp, 1)
    loss = (loss_fn(Xp[:, :, 0], xb[:, :, 0]) + torch.mean(1 - torch.cos(Xp[:, :, 1] - xb[:, :, 1])) +
            loss_fn(Xp[:, :, 2:], xb[:, :, 2:]))
    grads_of(model, loss, "phnn_euler")
    out["phnn_euler/x0_grad"] = x0b.grad.numpy().copy()
    out["phnn_euler/traj"] = Xp.detach().numpy()
    for k, v in model.state_dict().items():
        out["sd_phnn/" + k] = v.detach().numpy().copy()
    # RK4 trajectory matching through the reference's differentiable rollout
    Ug = ub[:, :-1, :].clone().requires_grad_(True)
    tr = rollout_trajectory_differentiable(model, xb[:, 0, :].clone().requires_grad_(True), Ug, dt, "rk4")
    loss = ((tr - xb) ** 2).mean()
    grads_of(model, loss, "phnn_rk4")
    out["phnn_rk4/U_grad"] = Ug.grad.numpy().copy()

    # --- canonical (scripts/train_cartpole_phnn_canonical.py:83-180, integrator='euler') ---
    torch.manual_seed(0)
    cm = pHNN_Canonical(os.path.join(CFG, "cartpole_phnn.yaml"))
    with torch.no_grad():
        cm.M_net.log_a.copy_(torch.tensor(0.25)); cm.M_net.b.copy_(torch.tensor(0.35)); cm.M_net.log_c.copy_(torch.tensor(-0.4))
        cm.R_diag_raw.copy_(torch.tensor([0.1, -0.3, 0.5, 1.2]))
    y0 = xb[:, 0, :].clone().requires_grad_(True)
    ys, vel = [y0], []
    for t in range(xb.shape[1] - 1):
        dy, _, inter = cm(ys[-1], ub[:, t, :], return_intermediate=True)
        ys.append(ys[-1] + dt * dy)
        _, qd_true = split_state(xb[:, t, :])
        vel.append(torch.sum((inter["q_dot_reconstructed"] - qd_true) ** 2, dim=1).mean())
    yp = torch.stack(ys, 1)
    l_pos = torch.mean((yp[:, :, 0] - xb[:, :, 0]) ** 2) + torch.mean(1 - torch.cos(yp[:, :, 1] - xb[:, :, 1]))
    loss = l_pos + torch.mean(torch.stack(vel))
    grads_of(cm, loss, "canon_euler")
    out["canon_euler/traj"] = yp.detach().numpy()
    for k, v in cm.state_dict().items():
        out["sd_canon/" + k] = v.detach().numpy().copy()
    np.savez(os.path.join(HERE, "train_grads.npz"), **out)
    print("train_grads done; losses", out["phnn_euler/loss"], out["phnn_rk4/loss"], out["canon_euler/loss"])


if __name__ == "__main__":
    which = sys.argv[1:] or ["pendulum", "h128", "h256", "canonical", "closed_loop", "cfg4_shape", "cfg5_shape", "train", "canonical_constM", "dropout", "nobias", "lbfgs"]
    if "pendulum" in which:
        gen_pendulum()
    if "h128" in which:
        gen_cartpole(128, 0, "cartpole_h128")
    if "h256" in which:
        gen_cartpole(256, 1234, "cartpole_h256")
    if "canonical" in which:
        gen_canonical()
    if "closed_loop" in which:
        gen_closed_loop()
    if "cfg4_shape" in which:
        gen_cfg4_shape()
    if "cfg5_shape" in which:
        gen_cfg5_shape()
    if "canonical_constM" in which:
        gen_canonical_constM()
    if "dropout" in which:
        gen_cartpole_dropout()
    if "nobias" in which:
        gen_canonical_nobias()
    if "lbfgs" in which:
        gen_lbfgs()
    if "train" in which:
        gen_train()
