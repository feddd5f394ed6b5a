"""Batched closed loop on the device: B cart-pole plants, each driven by its own gradient MPC.

One simulation step is what the reference's drivers do per instance in Python
(scripts/run_cartpole_mpc.py:121-176, scripts/run_mpc_canonical.py:55-95):
    controller(state) -> first control -> log H -> stability bookkeeping -> CartPoleSimulator.step
Here every piece is a kernel on one stream (cast, fused MPC solve, forward for H, plant step,
warm-start shift); the host only enqueues launches and never reads device data inside the loop.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from .batched import BatchedMPC


class ClosedLoopBatch:
    def __init__(self, mpc: BatchedMPC, dt_plant=None, warm_start=False, target=None, tolerance=None, min_duration=0.2,
                 log_energy=True):
        """mpc: the per-step solver (cold start each step when warm_start=False, as MPCController; shifted previous
        plan when True, as MPCControllerCanonical.control).  tolerance / min_duration: the stability criterion of
        the reference configs (`stability.tolerance`, `stability.min_duration`)."""
        self.mpc = mpc
        self.dt = float(mpc.dt if dt_plant is None else dt_plant)
        self.warm_start = bool(warm_start)
        self.target = np.zeros(4) if target is None else np.asarray(target, np.float64).reshape(4)
        self.tol = np.full(4, np.inf) if tolerance is None else np.asarray(tolerance, np.float64).reshape(4)
        self.min_duration = float(min_duration)
        self.log_energy = bool(log_energy)

    def run(self, initial_states, steps):
        """initial_states [B,4] (float64); returns a dict of device tensors: states [B,steps+1,4] f64, controls
        [B,steps] f32, energies [B,steps] f32 (H at the controller's state), done_step [B] (-1 = ran to the end),
        stability_achieved [B], stable_duration [B]."""
        L = _lib.lib()
        pk = self.mpc.pack
        dev = pk.device
        s0 = torch.as_tensor(np.asarray(initial_states, np.float64) if not isinstance(initial_states, torch.Tensor)
                             else initial_states).to(device=dev, dtype=torch.float64).reshape(-1, 4).contiguous()
        B, H = s0.shape[0], self.mpc.horizon
        state = s0.clone()
        traj = torch.empty((B, steps + 1, 4), dtype=torch.float64, device=dev)
        controls = torch.zeros((B, steps), dtype=torch.float32, device=dev)
        energies = torch.zeros((B, steps), dtype=torch.float32, device=dev)
        done = torch.full((B,), -1, dtype=torch.int32, device=dev)
        sstart = torch.full((B,), -1, dtype=torch.int32, device=dev)
        sdur = torch.zeros((B,), dtype=torch.float32, device=dev)
        sach = torch.zeros((B,), dtype=torch.int32, device=dev)
        ep = _lib.Episode(state.data_ptr(), traj.data_ptr(), controls.data_ptr(), done.data_ptr(), sstart.data_ptr(),
                          sdur.data_ptr(), sach.data_ptr(), int(steps))
        x0 = torch.empty((B, 4), dtype=torch.float32, device=dev)
        Uwarm = None
        tgt = self.target.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        tol = self.tol.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for k in range(steps):
                _lib.check(L.phnn_state_to_f32(state.data_ptr(), x0.data_ptr(), traj.data_ptr() if k == 0 else None,
                                               int(steps), B, stream), "phnn_state_to_f32")
                out = self.mpc.solve(x0, Uwarm)
                U = out["U"]
                if self.log_energy:
                    _, Hk = ops.forward(pk.handle, x0, U[:, 0].contiguous())
                    energies[:, k] = Hk
                _lib.check(L.phnn_plant_step(ctypes.byref(ep), U.data_ptr(), H, k, self.dt, tgt, tol, self.min_duration, B,
                                             stream), "phnn_plant_step")
                if self.warm_start:
                    Uwarm = torch.empty_like(U)
                    _lib.check(L.phnn_shift_controls(U.data_ptr(), Uwarm.data_ptr(), B, H, stream), "phnn_shift_controls")
        return {"states": traj, "controls": controls, "energies": energies, "done_step": done,
                "stability_achieved": sach.bool(), "stable_duration": sdur}
