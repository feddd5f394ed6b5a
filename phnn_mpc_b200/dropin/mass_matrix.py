"""Cart-pole mass matrix M(theta) = [[a, b cos], [b cos, c]] with a = exp(log_a)+1e-3,
c = exp(log_c)+1e-3 (src/mass_matrix.py:239-362).  Host-side helper: inside the MPC path the
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
