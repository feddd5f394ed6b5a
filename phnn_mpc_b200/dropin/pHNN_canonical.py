"""Drop-in for src/pHNN_canonical.py: ``forward(y, u, return_intermediate=False) ->
(dy [B,n], H [B], dict|None)``, state_dict keys R_diag_raw, J, G, M_net.{log_a,b,log_c}, H_net.*."""
import torch
import torch.nn as nn
import yaml

try:
    from .NN import MLP
    from .mass_matrix import CartPoleMassMatrix, MassMatrixNetwork
    from .coordinate_transforms import kinematic_to_canonical, momentum_to_velocity, split_state
    from .pHNN import _mlp_from_config, run_forward_op, _flatten_rows
except ImportError:
    from NN import MLP  # noqa: F401
    from mass_matrix import CartPoleMassMatrix, MassMatrixNetwork
    from coordinate_transforms import kinematic_to_canonical, momentum_to_velocity, split_state
    from pHNN import _mlp_from_config, run_forward_op, _flatten_rows


class pHNN_Canonical(nn.Module):
    def __init__(self, config_path: str):
        super().__init__()
        with open(config_path, "r") as f:
            model_cfg = yaml.safe_load(f)["model"]
        self.state_dim = model_cfg["state_dim"]
        self.input_dim = model_cfg["input_dim"]
        self.q_dim = self.state_dim // 2
        mm = model_cfg.get("mass_matrix", {})
        if mm.get("type", "cartpole") == "cartpole":
            self.M_net = CartPoleMassMatrix(init_a=mm.get("init_a", 1.0), init_b=mm.get("init_b", 0.1),
                                            init_c=mm.get("init_c", 1.0))
        else:
            # src/pHNN_canonical.py:79-86: any other type is a MassMatrixNetwork; 'constant' is built for the CUDA path,
            # the configuration-dependent types ('diagonal', 'full') raise NotImplementedError
            if self.q_dim != 2:
                raise NotImplementedError("the CUDA kernels are built for q_dim = 2")
            self.M_net = MassMatrixNetwork(q_dim=self.q_dim, mass_type=mm["type"],
                                           hidden_sizes=mm.get("hidden_sizes", [64, 64]),
                                           init_scale=mm.get("init_scale", 1.0))
        self.H_net = _mlp_from_config(model_cfg["H_mlp"], self.state_dim, 1)
        if not self.H_net.kernel_compatible():
            raise NotImplementedError("CUDA kernels cover Tanh MLPs without LayerNorm")
        qd = self.q_dim
        J = torch.zeros(2 * qd, 2 * qd)
        J[:qd, qd:] = torch.eye(qd)
        J[qd:, :qd] = -torch.eye(qd)
        self.register_buffer("J", J)
        self.R_diag_raw = nn.Parameter(torch.full((self.state_dim,), 0.1))
        if not model_cfg.get("fixed_G", False):
            raise ValueError("pHNN_Canonical requires fixed_G=True")
        self.register_buffer("G", torch.tensor(model_cfg["G_value"], dtype=torch.float32))

    def get_R_matrix(self, batch_size: int):
        r = torch.nn.functional.softplus(self.R_diag_raw) + 1e-4
        return torch.diag(r).unsqueeze(0).expand(batch_size, -1, -1)

    def get_velocity_reconstruction(self, y):
        z = kinematic_to_canonical(y, self.M_net)
        q, p = split_state(z)
        return momentum_to_velocity(q, p, self.M_net)

    def forward(self, y, u, return_intermediate: bool = False):
        dy, H = run_forward_op(self, y, u)
        if not return_intermediate:
            return dy, H, None
        # diagnostics only (host-side, detached): the quantities the reference exposes
        with torch.no_grad():
            y2 = _flatten_rows(y).detach().float().cpu()
            z = kinematic_to_canonical(y2, self.M_net)
            q, p = split_state(z)
            inter = {"z": z, "q": q, "p": p, "q_dot_reconstructed": dy[:, : self.q_dim].detach(),
                     "R": self.get_R_matrix(y2.shape[0])}
        return dy, H, inter
