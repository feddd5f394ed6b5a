"""phnn_mpc_b200 -- B200-native (sm_100a) implementation of the pHNN-MPC hot path.

Layout
  csrc/        hand-written CUDA kernel + the C ABI declared in include/phnn_mpc.h
  _lib.py      ctypes binding of libphnn_mpc.so (no fallback when it is missing)
  packing.py   reference state_dict -> packed device weights
  ops.py       torch.library custom ops ``phnn_mpc::{forward,vjp,rollout,cost_grad,mpc_solve}``
  batched.py   batched front-end (B instances per call) used by bench.py and the drop-ins
  dropin/      modules with the reference's names and signatures (pHNN, pHNN_Canonical,
               integrators, mpc_controller, mpc_controller_canonical)
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
