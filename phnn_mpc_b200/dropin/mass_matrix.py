"""Cart-pole mass matrix M(theta) = [[a, b cos], [b cos, c]] with a = exp(log_a)+1e-3,
c = exp(log_c)+1e-3 (src/mass_matrix.py:239-362), and the constant variant of MassMatrixNetwork
(src/mass_matrix.py:15-216, mass_type='constant').  Host-side helpers: inside the MPC path the
same formulas are evaluated by the CUDA kernel (csrc/phnn_kernel.cuh, canon_of)."""
import torch
import torch.nn as nn


class CartPoleMassMatrix(nn.Module):
    def __init__(self, init_a: float = 1.0, init_b: float = 0.1, init_c: float = 1.0):
        super().__init__()
        self.log_a = nn.Parameter(torch.log(torch.tensor(init_a)))
        self.b = nn.Parameter(torch.tensor(init_b))
        self.log_c = nn.Parameter(torch.log(torch.tensor(init_c)))

    def _abc(self):
        # the reference detaches these through .item(); only cos(theta) carries gradient
        return (float(torch.exp(self.log_a) + 1e-3), float(self.b), float(torch.exp(self.log_c) + 1e-3))

    def forward(self, q):
        a, b, c = self._abc()
        bc = b * torch.cos(q[:, 1])
        row0 = torch.stack([torch.full_like(bc, a), bc], dim=1)
        row1 = torch.stack([bc, torch.full_like(bc, c)], dim=1)
        return torch.stack([row0, row1], dim=1)

    def inverse(self, q):
        a, b, c = self._abc()
        bc = b * torch.cos(q[:, 1])
        det = (torch.full_like(bc, a) * c - bc ** 2) + 1e-6
        row0 = torch.stack([torch.full_like(bc, c) / det, -bc / det], dim=1)
        row1 = torch.stack([-bc / det, torch.full_like(bc, a) / det], dim=1)
        return torch.stack([row0, row1], dim=1)

    def get_parameters_dict(self):
        a, b, c = self._abc()
        return {"a": a, "b": b, "c": c}


class MassMatrixNetwork(nn.Module):
    """Drop-in for src/mass_matrix.py:15-216, mass_type 'constant' only: M = L L^T with L = tril(L_tril) and
    softplus(diag) + 1e-3 on the diagonal, independent of q (state_dict key ``L_tril``).  The configuration-dependent
    types ('diagonal', 'full': an MLP of q) are not built for the CUDA path and raise."""

    def __init__(self, q_dim: int, mass_type: str = "diagonal", hidden_sizes=(64, 64), activation=None,
                 init_scale: float = 1.0):
        super().__init__()
        if mass_type not in ("constant", "diagonal", "full"):
            raise ValueError(f"Unknown mass_type: {mass_type}")
        if mass_type != "constant":
            raise NotImplementedError("MassMatrixNetwork mass_type=%r is not built for the CUDA path; 'constant' and the "
                                      "cart-pole mass matrix are" % (mass_type,))
        self.q_dim, self.mass_type, self.init_scale = q_dim, mass_type, init_scale
        self.L_tril = nn.Parameter(torch.eye(q_dim) * init_scale)
        self.mlp = None

    def _L(self):
        L = torch.tril(self.L_tril).clone()
        idx = torch.arange(self.q_dim)
        L[idx, idx] = torch.nn.functional.softplus(L[idx, idx]) + 1e-3
        return L

    def forward(self, q):
        L = self._L()
        return (L @ L.T).unsqueeze(0).expand(q.shape[0], -1, -1)

    def inverse(self, q):
        Li = torch.inverse(self._L())
        return (Li.T @ Li).unsqueeze(0).expand(q.shape[0], -1, -1)
