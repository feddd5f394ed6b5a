"""y = [q, qdot] <-> z = [q, p] with p = M(q) qdot (src/coordinate_transforms.py:20-130).
Host-side helpers; the MPC path does these transforms in the CUDA kernel."""
import torch


def split_state(state):
    half = state.shape[1] // 2
    return state[:, :half], state[:, half:]


def velocity_to_momentum(q, q_dot, M_net):
    return torch.bmm(M_net(q), q_dot.unsqueeze(-1)).squeeze(-1)


def momentum_to_velocity(q, p, M_net):
    return torch.bmm(M_net.inverse(q), p.unsqueeze(-1)).squeeze(-1)


def kinematic_to_canonical(y, M_net):
    q, q_dot = split_state(y)
    return torch.cat([q, velocity_to_momentum(q, q_dot, M_net)], dim=1)


def canonical_to_kinematic(z, M_net):
    q, p = split_state(z)
    return torch.cat([q, momentum_to_velocity(q, p, M_net)], dim=1)
