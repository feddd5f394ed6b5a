#!/usr/bin/env python
"""bench.py -- headline benchmark of the pHNN-MPC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one complete batched MPC solve (iters x {clamp, rollout, cost, adjoint, Adam}) of
the workload's B instances per GPU; N GPUs = N independent shards (weak scaling) + one final
NCCL gather of the controls.  Prints ONE JSON line (see the task contract); `value` is
whole-job solves/s with inputs resident in HBM, `e2e` the same through the public Python API
with host buffers (H2D of x0, D2H of controls + cost inside the timed region).

--impl reference times the CPU implementation of the same path (the oracle port: the reference
is pure Python/PyTorch and its sources cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# BASELINE.json configs; FLOP model from SURVEY.md section 8(d) (1 FMA = 2 FLOP, no recompute)
WORKLOADS = {
    # name: (golden fixture, kind, B per GPU, H, iters, integrator, lr, description)
    "cfg4_rk4": ("cartpole_h256", "phnn", 65536, 50, 20, "rk4", 0.015,
                 "BASELINE cfg4: 65536 cart-pole MPC instances x H=50 x 20 Adam iters, RK4, pHNN hidden 256 "
                 "(random init seed 1234)"),
    "cfg4_euler": ("cartpole_h256", "phnn", 65536, 50, 20, "euler", 0.015,
                   "BASELINE cfg4 (Euler variant): 65536 instances x H=50 x 20 iters, pHNN hidden 256"),
    "cfg3": ("canonical", "canonical", 1024, 10, 50, "euler", 0.03,
             "BASELINE cfg3: canonical pHNN pole-stabilisation MPC, 1024 instances x H=10 x 50 iters, Euler"),
    "cfg5_h100": ("cartpole_h256", "phnn", 131072, 100, 20, "rk4", 0.015,
                  "BASELINE cfg5 shard: 1M/8 instances x H=100 x 20 iters, RK4, hidden 256"),
    "small": ("cartpole_h256", "phnn", 9472, 50, 2, "rk4", 0.015,
              "profiling-sized cfg4 slice (one full wave, 64 instances per SM): 9472 instances x H=50 x 2 iters, "
              "RK4, hidden 256"),
}


def flops_per_eval(kind, h, n):
    if kind == "canonical":
        # H fwd 2(4h+h^2+h) + dH 2(h^2+4h) + small ; HVP 4(h^2+4h) + small  (67 880 / 67 664 @128)
        return 4 * h * h + 18 * h + 40, 4 * h * h + 16 * h + 80
    return 4 * h * h + 58 * h + 200, 4 * h * h + 56 * h + 400


def solve_flops(kind, h, n, H, iters, S):
    ff, fv = flops_per_eval(kind, h, n)
    return iters * H * S * (ff + fv)


def load_fixture(name):
    z = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    return {k[3:]: z[k] for k in z.files if k.startswith("sd/")}


def cost_for(kind):
    if kind == "canonical":   # pole_stabilization.yaml
        return dict(Q=[0.0, 1000.0, 0.0, 100.0], R=[1e-4], u_min=-30.0, u_max=30.0)
    return dict(Q=[10.0, 200.0, 1.0, 10.0], R=[0.01], u_min=-15.0, u_max=15.0)


def make_inputs(B, kind, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    x0 = (torch.rand(B, 4, generator=g) * 2 - 1) * torch.tensor([1.0, 0.3, 0.5, 0.5])
    if kind == "canonical":
        x0 = x0 + torch.tensor([0.0, 0.05, 0.0, 0.0])
    return x0.float().contiguous()


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons during the timed region (pynvml, else nvidia-smi)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.2)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_leg(args, wl, steps, warmup, threads=None, budget_s=20.0):
    """times the oracle port (CPU restatement of the reference path) on a bounded sample"""
    from oracle.phnn_oracle import OracleModel, set_threads, max_threads
    fixture, kind, B, H, iters, integ, lr, desc = wl
    sd = load_fixture(fixture)
    M = OracleModel(sd, kind)
    cores = threads or os.cpu_count() or max_threads()
    set_threads(cores)
    c = cost_for(kind)
    C = M.cost_struct(c["Q"], c["R"], np.zeros(4), c["u_min"], c["u_max"])
    x0 = make_inputs(max(cores * 8, 64), kind, 7).numpy()
    # calibrate the sample: one instance per thread, one iteration
    nb = cores
    t0 = time.perf_counter()
    M.mpc_solve(C, x0[:nb], np.zeros((nb, H, 1), np.float32), 0.02, integ, lr=lr, iters=1,
                return_mode="last" if kind == "phnn" else "best")
    t1 = time.perf_counter() - t0
    per_round = t1 * iters                      # seconds for `cores` full solves
    rounds = max(1, int(budget_s / max(per_round, 1e-3) / max(1, steps + warmup)))
    nb = min(x0.shape[0], cores * rounds)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        M.mpc_solve(C, x0[:nb], np.zeros((nb, H, 1), np.float32), 0.02, integ, lr=lr, iters=iters,
                    return_mode="last" if kind == "phnn" else "best")
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return {"value": nb / (ms / 1e3), "unit": "solves/s", "cores": int(cores), "kind": "port",
            "sample": "%d of the workload's instances per step (full H=%d, %d iters, %s), %d step(s)" % (
                nb, H, iters, integ, len(times)), "ms_per_step": ms, "B_sample": int(nb)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4_rk4", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tensor-mode", type=int, default=-1, help="0 FP32-FMA kernel, 2 tcgen05 TF32 + BF16 correction product (default), 3 tcgen05 3xTF32, 1 tcgen05 plain TF32")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.batch:
        wl[2] = args.batch
    fixture, kind, B, H, iters, integ, lr, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    S = 4 if integ == "rk4" else 1
    sd = load_fixture(fixture)
    h = int(sd["H_net.net.0.weight"].shape[0])
    config = {"workload": desc, "name": args.workload, "instances_per_gpu": B, "horizon": H, "iters": iters,
              "integrator": integ, "hidden": h, "model_kind": kind, "parallelism": "instance-sharded x%d" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        warm = max(1, min(args.warmup, 1))
        res = cpu_leg(args, wl, args.steps, warm, budget_s=60.0)
        line = {"impl": "reference", "metric": "cartpole_mpc_solves_per_s", "value": res["value"], "unit": "solves/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warm, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": res["value"], "unit": "solves/s", "cores": res["cores"], "kind": "port",
                                 "sample": res["sample"]},
                "e2e": {"value": res["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from phnn_mpc_b200 import _lib
    from phnn_mpc_b200.batched import BatchedMPC, CostSpec
    from phnn_mpc_b200.packing import PackedModel
    from phnn_mpc_b200.distributed import gather_shards

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, kind, device=dev)
    if args.tensor_mode >= 0:
        pk.set_option("tensor_mode", args.tensor_mode)
    c = cost_for(kind)
    spec = CostSpec.make(4, 1, c["Q"], c["R"], None, c["u_min"], c["u_max"])
    mpc = BatchedMPC(pk, H, 0.02, spec, integrator=integ, lr=lr, iters=iters,
                     return_mode="last" if kind == "phnn" else "best")
    x0_host = make_inputs(B, kind, 7 + rank).pin_memory()
    x0 = x0_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed between ----
    for _ in range(args.warmup):
        out = mpc.solve(x0)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = mpc.solve(x0)
        e1.record()
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    # final gather of the controls (the path's only collective); timed separately, reported
    gather_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        Uall = gather_shards(out["U"], B * world)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
        assert Uall.shape[0] == B * world
        total_ms += gather_ms * args.steps
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        U_host = torch.empty((B, H, 1), dtype=torch.float32).pin_memory()
        c_host = torch.empty((B,), dtype=torch.float32).pin_memory()
        n_e2e = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            xd = x0_host.to(dev, non_blocking=True)
            o = mpc.solve(xd)
            U_host.copy_(o["U"], non_blocking=True)
            c_host.copy_(o["best_cost"], non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world / (float(t.item()) / 1e3), "unit": "solves/s",
               "h2d_bytes_per_step": int(x0_host.numel() * 4), "d2h_bytes_per_step": int((U_host.numel() + B) * 4),
               "steps": n_e2e}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the solve kernel ----
    # achieved = ALGORITHMIC FLOPs of the launch (iters*H*S*(F_f+F_vjp) per instance, no
    # recompute, 1 FMA = 2 FLOP) / CUDA-event duration.  The tcgen05 kernel is judged against the measured
    # dense bf16 tensor peak (MEASURED_PEAKS.json); the FP32-FMA kernel against the FP32 rate measured here.
    import ctypes
    L = _lib.lib()
    probe_out = torch.zeros(4, device=dev)
    fl = ctypes.c_double()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        L.phnn_ffma_probe(ctypes.c_void_p(probe_out.data_ptr()), 4000, 148 * 8, st, ctypes.byref(fl))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.phnn_ffma_probe(ctypes.c_void_p(probe_out.data_ptr()), 20000, 148 * 8, st, ctypes.byref(fl))
    e1.record()
    torch.cuda.synchronize()
    fp32_peak = fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    L.phnn_tf32_probe(ctypes.c_void_p(probe_out.data_ptr()), 2048, 148, st, ctypes.byref(fl))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.phnn_tf32_probe(ctypes.c_void_p(probe_out.data_ptr()), 16384, 148, st, ctypes.byref(fl))
    e1.record()
    torch.cuda.synchronize()
    tf32_peak = fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    algo = solve_flops(kind, h, 4, H, iters, S) * B
    kernel_ms = float(np.mean(step_ms))
    achieved = algo / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tmode = pk.get_option("tensor_mode")
    uses_tc = tmode in (1, 2, 3) and B >= pk.get_option("tensor_min_batch")
    traffic = None
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        if tr.get("workload") == args.workload and tr.get("B") == B:
            traffic = tr.get("dram_bytes_per_launch")
        elif tr.get("h") == h and tr.get("dram_bytes_per_tile_eval_pair") and kind == "phnn":
            # the capture is of a slice of the same job: DRAM traffic is proportional to (tiles x evaluation pairs)
            traffic = tr["dram_bytes_per_tile_eval_pair"] * ((B + 127) // 128) * iters * H * S
    except Exception:
        pass
    if uses_tc:
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        # tensor work actually issued: per evaluation 2 (forward) + 2 (adjoint: the two Hessian-vector products; the
        # forward activations come back from the tape) products of 128 x h x h MACs per 128-instance tile.  Mode 3
        # issues each as 3 TF32 MMAs; mode 2 as 1 TF32 MMA + one BF16 product of twice the depth ([a_lo | a] x [b | b_lo]);
        # mode 1 as 1 TF32 MMA.  BF16 MMAs run at twice the TF32 rate, so the time at peak rate is counted in TF32 units.
        tiles = (B + 127) // 128
        base = tiles * iters * H * S * (2 + 2) * 2.0 * 128 * h * h
        tf32_flops = base * (3 if tmode == 3 else 1)
        bf16_flops = base * 2 if tmode == 2 else 0.0
        mma_flops = tf32_flops + bf16_flops
        tf32_equiv_flops = tf32_flops + bf16_flops / 2
        scheme = {3: "3xTF32 error-compensated", 2: "TF32 + BF16 correction product (FP32-level accuracy)", 1: "plain TF32"}[tmode]
        # HBM side of the same kernel: the activation tape (a1, a2, g1: 3 h floats per instance and evaluation) is
        # written by the forward sweep and read back by the adjoint
        tape_bytes = tiles * iters * H * S * 2.0 * 3 * h * 128 * 4
        hbm_peak = peaks.get("hbm_gbs", 6500.0) if peaks else 6500.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": achieved / bf16_peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, this pool)" if peaks else
                                   "fallback 1400 (B200_PROFILING.md)",
                    "algorithmic_flops_per_launch": algo,
                    "kernel": "phnn_tc_kernel<MK,NS,HID> (tcgen05 kind::tf32%s, %s), one launch per step" % (
                        " + kind::f16" if tmode == 2 else "", scheme),
                    "kernel_ms": kernel_ms,
                    "executed_tensor_tflops": mma_flops / (kernel_ms * 1e-3) / 1e12,
                    "tf32_mma_peak_tflops": tf32_peak,
                    "tensor_pipe_frac": tf32_equiv_flops / (kernel_ms * 1e-3) / 1e12 / tf32_peak,
                    "executed_over_algorithmic": mma_flops / algo,
                    "fp32_fma_peak_tflops": fp32_peak, "frac_of_fp32_fma_peak": achieved / fp32_peak,
                    "hbm": {"algorithmic_tape_bytes_per_launch": tape_bytes,
                            "achieved_gbps": tape_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbps": hbm_peak,
                            "frac": tape_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                            "measured_dram_bytes_per_launch": traffic},
                    "note": "FP32-level accuracy on the tensor cores costs one TF32 product (half the bf16 rate) plus a BF16 "
                            "correction product of twice the depth: frac vs the bf16 peak is bounded by 1/4; the adjoint "
                            "reads the forward activations from an HBM tape instead of recomputing them (4 tensor products "
                            "per pair instead of 6); the binding resource is the L1/shared-memory data pipe (DESIGN.md 4a)"}
    else:
        roofline = {"bound": "fp32-fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak, "traffic": traffic,
                    "peak_source": "FFMA probe kernel timed in this run (phnn_ffma_probe); nominal 74.4 at 1965 MHz",
                    "algorithmic_flops_per_launch": algo,
                    "frac_of_measured_bf16_tensor_peak": (achieved / peaks["bf16_tflops_sustained"]) if peaks else None,
                    "kernel": "phnn_kernel<MK,NS,HID> (one launch per step)", "kernel_ms": kernel_ms}

    # ---- second half of BASELINE.json's metric: pHNN RK4 rollout steps/s (config 2), same run ----
    rollout_metric = None
    try:
        from phnn_mpc_b200.batched import rollout as _rollout
        sdp = load_fixture("pendulum")
        pkp = PackedModel({k: torch.from_numpy(v) for k, v in sdp.items()}, "phnn", device=dev)
        g = torch.Generator().manual_seed(1)
        Bp, Tp = 4096, 100
        xp = torch.stack([(torch.rand(Bp, generator=g) * 2 - 1) * np.pi, torch.rand(Bp, generator=g) * 2 - 1], 1).to(dev)
        Up = (torch.rand(Bp, Tp, 1, generator=g) * 4 - 2).to(dev)
        for _ in range(3):
            _rollout(pkp, xp, Up, 0.05, "rk4")
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _rollout(pkp, xp, Up, 0.05, "rk4")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        rms = float(np.median(ts))
        rollout_metric = {"metric": "phnn_rk4_rollout_instance_steps_per_s", "value": Bp * Tp / (rms * 1e-3), "unit": "instance-steps/s",
                          "ms": rms, "config": "BASELINE cfg2: pendulum pHNN (shipped weights, h=64, learned G), 4096 initial states x H=100, "
                                               "RK4, dt 0.05, one launch (FP32-FMA kernel)",
                          "algorithmic_tflops": Bp * Tp * 73360 / (rms * 1e-3) / 1e12}
    except Exception as ex:  # the headline line must still be printed
        rollout_metric = {"error": repr(ex)}

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_leg(args, wl, 1, 0, budget_s=15.0)
        cpu = {"value": r["value"], "unit": "solves/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {"metric": "cartpole_mpc_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, kernel_path={3: "tcgen05-3xTF32", 2: "tcgen05-TF32+BF16corr", 1: "tcgen05-TF32"}[tmode] if uses_tc else "fp32-fma",
                           l2="256 MiB buffer written between timed steps (L2 flush); per-step working set "
                                      "(stage checkpoints) also exceeds L2", gather_ms=gather_ms),
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu, "rollout": rollout_metric,
            "step_ms": step_ms}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
