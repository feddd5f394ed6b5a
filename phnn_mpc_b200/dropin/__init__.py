"""Drop-in modules with the reference's module names, class names and call signatures.

Put this directory ahead of the reference's ``src`` on ``sys.path`` (the reference's drivers
*append* ``src``, scripts/run_cartpole_mpc.py:21) and ``from pHNN import pHNN`` etc. resolve
here; ``cartpole_simulator`` and anything else not provided still comes from the reference.
The arithmetic of every forward / rollout / MPC solve runs in the CUDA library; these classes
only hold parameters (same state_dict keys) and stage tensors.
"""
