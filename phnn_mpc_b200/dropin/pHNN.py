"""Drop-in for src/pHNN.py: same constructor, parameters (state_dict keys J, G_fixed |
G_net.*, R_net.*, H_net.*) and ``forward(x, u) -> (dx [B,n], H [B])``; the evaluation runs in
the CUDA library (op phnn_mpc::forward) instead of ~25 ATen calls plus an autograd tape."""
import torch
import torch.nn as nn
import yaml

try:
    from .NN import MLP
except ImportError:  # imported as a top-level module from sys.path, like the reference's src/
    from NN import MLP

from phnn_mpc_b200 import ops
from phnn_mpc_b200.packing import pack_of


def _mlp_from_config(spec, n_in, n_out):
    act = getattr(nn, spec["activation"].split(".")[-1])
    return MLP(input_dim=n_in, output_dim=n_out, hidden_sizes=tuple(spec["hidden_sizes"]), activation=act,
               dropout=spec["dropout"], layer_norm=spec["layer_norm"], bias=spec["bias"])


def _flatten_rows(t):
    """any rank -> [B, last] (src/pHNN.py:58-66)."""
    return t.unsqueeze(0) if t.ndim == 1 else t.reshape(-1, t.shape[-1])


def run_forward_op(module, x, u):
    """Stage (x, u) to the module's CUDA pack, run the op, return results on x's device."""
    x2, u2 = _flatten_rows(x), _flatten_rows(u)
    for net in (getattr(module, "H_net", None), getattr(module, "R_net", None), getattr(module, "G_net", None)):
        if net is not None and hasattr(net, "check_mode"):
            net.check_mode()
    if not torch.cuda.is_available():
        raise RuntimeError("phnn_mpc_b200 has no CPU fallback: a CUDA device is required for forward()")
    pk = pack_of(module)
    xd = x2.to(device=pk.device, dtype=torch.float32)
    ud = u2.to(device=pk.device, dtype=torch.float32)
    dx, H = ops.forward(pk.handle, xd, ud)
    return dx.to(x2.device), H.detach().to(x2.device)


class pHNN(nn.Module):
    def __init__(self, config_path: str):
        super().__init__()
        with open(config_path, "r") as f:
            model_cfg = yaml.safe_load(f)["model"]
        n, m = model_cfg["state_dim"], model_cfg["input_dim"]
        self.state_dim, self.input_dim = n, m
        self.J = nn.Parameter(torch.randn(n, n))  # used as J - J^T (no 1/2), src/pHNN.py:83
        self.R_net = _mlp_from_config(model_cfg["R_mlp"], n, n * n)
        self.H_net = _mlp_from_config(model_cfg["H_mlp"], n, 1)
        if model_cfg.get("fixed_G", False):
            self.register_buffer("G_fixed", torch.tensor(model_cfg["G_value"], dtype=torch.float32))
            self.G_net = None
        else:
            self.G_net = _mlp_from_config(model_cfg["G_mlp"], n, m * n)
        nets = [self.R_net, self.H_net] + ([self.G_net] if self.G_net is not None else [])
        if not all(net.kernel_compatible() for net in nets):
            raise NotImplementedError("CUDA kernels cover Tanh MLPs without LayerNorm "
                                      "(all shipped configs); got another variant")

    def forward(self, x, u):
        return run_forward_op(self, x, u)
