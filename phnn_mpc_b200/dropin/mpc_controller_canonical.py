"""Drop-in for src/mpc_controller_canonical.py: same constructor, ``control(x, u_prev) ->
(u [m], info)`` with warm start, best-iterate tracking, cost history and solve_time; the
optimisation loop (src/mpc_controller_canonical.py:163-228) is one CUDA launch."""
import time

import numpy as np
import torch

from phnn_mpc_b200.batched import BatchedMPC, CostSpec, rollout as _rollout


class MPCControllerCanonical:
    def __init__(self, model, horizon=20, dt=0.02, Q=None, R=None, x_target=None, u_min=-10.0, u_max=10.0,
                 optimizer_steps=50, learning_rate=0.1, verbose=False):
        self.model = model
        self.model.eval()
        self.horizon, self.dt = horizon, dt
        self.optimizer_steps, self.learning_rate, self.verbose = optimizer_steps, learning_rate, verbose
        self.state_dim, self.input_dim = model.state_dim, model.input_dim
        if Q is None:
            Q = np.diag([10.0, 100.0, 1.0, 10.0][: self.state_dim] + [1.0] * max(0, self.state_dim - 4))
        if R is None:
            R = 0.01 * np.eye(self.input_dim)
        if x_target is None:
            x_target = np.zeros(self.state_dim)
        self.Q = torch.tensor(Q, dtype=torch.float32)
        self.R = torch.tensor(R, dtype=torch.float32)
        self.x_target = torch.tensor(x_target, dtype=torch.float32)
        self.u_min, self.u_max = u_min, u_max

    def _engine(self, integrator="euler"):
        spec = CostSpec.make(self.state_dim, self.input_dim, self.Q, self.R, self.x_target, self.u_min, self.u_max)
        return BatchedMPC(self.model, self.horizon, self.dt, spec, integrator=integrator, lr=self.learning_rate,
                          iters=self.optimizer_steps, return_mode="best")

    def compute_cost(self, x_pred, u_seq):
        """src/mpc_controller_canonical.py:91-120"""
        e = x_pred - self.x_target
        return torch.einsum("ti,ij,tj->", e, self.Q, e) + torch.einsum("ti,ij,tj->", u_seq, self.R, u_seq)

    def rollout(self, x0, u_seq):
        """Euler rollout [H+1, n] (src/mpc_controller_canonical.py:122-161)"""
        traj = _rollout(self.model, x0.reshape(1, -1), u_seq.reshape(1, -1, self.input_dim), self.dt, "euler")
        return traj[0].to(x0.device)

    def solve_batch(self, states, U0=None, integrator="euler", want_hist=False):
        """B independent solves in one launch (best-iterate semantics per instance)."""
        return self._engine(integrator).solve(states, U0, want_hist)

    def optimize_control(self, x0, u_init=None):
        """(u_opt [H,m], info) -- best pre-step clamped iterate (src/mpc_controller_canonical.py:163-228)"""
        out = self._engine().solve(x0.reshape(1, -1), None if u_init is None else u_init.reshape(1, self.horizon, -1),
                                   want_hist=True)
        costs = out["cost_hist"][:, 0].cpu().tolist()
        if self.verbose:
            for step, c in enumerate(costs):
                if step % 10 == 0 or step == self.optimizer_steps - 1:
                    print(f"  Step {step:3d}: cost = {c:.4f}")
        info = {"costs": costs, "final_cost": float(out["best_cost"][0].cpu()), "num_steps": self.optimizer_steps}
        return out["U"][0].cpu(), info

    def control(self, x_current, u_prev=None):
        """receding-horizon action with shifted warm start (src/mpc_controller_canonical.py:230-273)"""
        t0 = time.time()
        x0 = torch.tensor(x_current, dtype=torch.float32)
        u_init = None
        if u_prev is not None:
            u_init = torch.tensor(u_prev, dtype=torch.float32)
            u_init = torch.cat([u_init[1:], torch.zeros(1, self.input_dim)], dim=0)
        u_opt, opt_info = self.optimize_control(x0, u_init)
        info = {"u_sequence": u_opt.numpy(), "solve_time": time.time() - t0, "optimization": opt_info}
        return u_opt[0].numpy(), info


def create_mpc_controller(model, config):
    """same schema as src/mpc_controller_canonical.py:276-316 (reads config['mpc'] and config['cartpole']['dt'])"""
    c = config.get("mpc", {})
    return MPCControllerCanonical(
        model=model, horizon=c.get("horizon", 20), dt=config["cartpole"]["dt"],
        Q=np.diag(c.get("Q_diag", [10.0, 100.0, 1.0, 10.0])), R=np.diag(c.get("R_diag", [0.01])),
        x_target=np.array(c.get("x_target", [0.0, 0.0, 0.0, 0.0])), u_min=c.get("u_min", -10.0),
        u_max=c.get("u_max", 10.0), optimizer_steps=c.get("optimizer_steps", 50),
        learning_rate=c.get("learning_rate", 0.1), verbose=c.get("verbose", False))
