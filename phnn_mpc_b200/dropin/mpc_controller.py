"""Drop-in for src/mpc_controller.py.  ``MPCController.compute_control`` keeps its signature and
return type (np.float32[1]) but the 30 x {clamp, 20-step Euler rollout, cost, backward, Adam}
loop (src/mpc_controller.py:143-209) is one launch of the fused CUDA solve kernel.
``solve_batch`` is the batched entry the reference lacks (B instances per launch)."""
import numpy as np
import torch

from phnn_mpc_b200 import ops
from phnn_mpc_b200.batched import BatchedMPC, CostSpec


class MPCController:
    def __init__(self, phnn_model, horizon, dt, Q, R, target_state=None, u_min=None, u_max=None, x_min=None,
                 x_max=None, optimizer_type="Adam", lr=0.1, max_iterations=50):
        self.model = phnn_model
        self.model.eval()
        self.horizon, self.dt = horizon, dt
        self.state_dim = phnn_model.J.shape[0]
        self.Q = torch.diag(torch.tensor(Q, dtype=torch.float32)) if isinstance(Q, list) else torch.diag(Q)
        self.R = R
        self.target_state = (torch.zeros(self.state_dim) if target_state is None
                             else torch.tensor(target_state, dtype=torch.float32))
        self.u_min, self.u_max = u_min, u_max
        self.x_min = torch.tensor(x_min, dtype=torch.float32) if x_min is not None else None
        self.x_max = torch.tensor(x_max, dtype=torch.float32) if x_max is not None else None
        self.optimizer_type, self.lr, self.max_iterations = optimizer_type, lr, max_iterations

    # -- pieces of the reference's public surface -------------------------------------------
    def _spec(self):
        return CostSpec.make(self.state_dim, 1, self.Q, float(self.R), self.target_state, self.u_min, self.u_max,
                             self.x_min, self.x_max)

    def _engine(self, integrator="euler"):
        return BatchedMPC(self.model, self.horizon, self.dt, self._spec(), integrator=integrator, lr=self.lr,
                          iters=self.max_iterations, return_mode="last")

    def compute_cost(self, states, controls):
        """quadratic tracking cost + control effort (+ soft state bounds), src/mpc_controller.py:75-114"""
        err = states - self.target_state
        total = torch.einsum("ti,ij,tj->", err, self.Q, err) + self.R * (controls ** 2).sum()
        if self.x_min is not None:
            total = total + 1000.0 * (torch.relu(self.x_min - states) ** 2).sum()
        if self.x_max is not None:
            total = total + 1000.0 * (torch.relu(states - self.x_max) ** 2).sum()
        return total

    def rollout_dynamics(self, x0, controls):
        """Euler rollout [H+1, n] of the model (src/mpc_controller.py:116-141)"""
        from phnn_mpc_b200.batched import rollout
        traj = rollout(self.model, x0.reshape(1, -1), controls.reshape(1, -1, 1), self.dt, "euler")
        return traj[0].to(x0.device)

    def _check_optimizer(self):
        if self.optimizer_type == "Adam":
            return
        if self.optimizer_type == "LBFGS":
            return
        raise ValueError(f"Unknown optimizer type: {self.optimizer_type}")

    def _compute_control_lbfgs(self, current_state, return_sequence=False):
        """src/mpc_controller.py:169-170,174-199: ``torch.optim.LBFGS(lr, max_iter=20)`` stepped ``max_iterations`` times.  The
        optimiser is torch's own (its two-loop recursion and stopping rules are the algorithm); its closure -- clamp, Euler
        rollout, cost, backward -- is one launch of the fused cost + adjoint kernel per evaluation."""
        eng = self._engine()
        x0 = current_state.reshape(1, -1)
        control_sequence = torch.zeros(self.horizon, 1, requires_grad=True)
        optimizer = torch.optim.LBFGS([control_sequence], lr=self.lr, max_iter=20)

        def closure():
            optimizer.zero_grad()
            cost, g, _ = eng.cost_and_grad(x0, control_sequence.detach().reshape(1, self.horizon, 1))
            control_sequence.grad = g[0].to("cpu", torch.float32).reshape(self.horizon, 1)
            return cost[0].to("cpu")

        for _ in range(self.max_iterations):
            optimizer.step(closure)
        with torch.no_grad():
            seq = control_sequence.detach()
            if self.u_min is not None and self.u_max is not None:
                seq = torch.clamp(seq, self.u_min, self.u_max)
        return seq.numpy() if return_sequence else seq[0].numpy()

    def solve_batch(self, states, U0=None, integrator="euler", want_hist=False):
        """B independent solves in one launch.  states [B,n] -> dict(U, u0, best_cost, cost_hist)."""
        self._check_optimizer()
        if self.optimizer_type != "Adam":
            raise NotImplementedError("solve_batch runs the fused Adam solve; the LBFGS branch is per instance (compute_control)")
        return self._engine(integrator).solve(states, U0, want_hist)

    def compute_control(self, current_state):
        """first control of the optimised sequence, cold start (src/mpc_controller.py:143-209)"""
        self._check_optimizer()
        if isinstance(current_state, np.ndarray):
            current_state = torch.tensor(current_state, dtype=torch.float32)
        if self.optimizer_type == "LBFGS":
            return self._compute_control_lbfgs(current_state)
        out = self._engine().solve(current_state.reshape(1, -1))
        return out["u0"][0].cpu().numpy()


def create_mpc_from_config(phnn_model, config):
    """same schema as src/mpc_controller.py:212-241 (mpc.{horizon,dt,Q,R,...})"""
    c = config["mpc"]
    return MPCController(phnn_model=phnn_model, horizon=c["horizon"], dt=c["dt"], Q=c["Q"], R=c["R"],
                         target_state=c.get("target_state"), u_min=c.get("u_min"), u_max=c.get("u_max"),
                         x_min=c.get("x_min"), x_max=c.get("x_max"), optimizer_type=c.get("optimizer", "Adam"),
                         lr=c.get("lr", 0.1), max_iterations=c.get("max_iterations", 50))
