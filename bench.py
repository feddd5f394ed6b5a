#!/usr/bin/env python
"""bench.py -- headline benchmark of the pHNN-MPC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one complete batched MPC solve (iters x {clamp, rollout, cost, adjoint, Adam}) of the
workload's instances on each GPU (N GPUs = N contiguous shards, no data-path collective) FOLLOWED BY
the path's only exchange, the final gather of the controls, inside the same timed bracket.
Prints ONE JSON line (see the task contract): `value` is whole-job solves/s with inputs resident in
HBM, `e2e` the same through the public Python API with host buffers (H2D of x0, D2H of controls +
cost inside the timed region, every step, L2 flushed between steps).

Workloads: cfg4_* keep 65 536 instances per GPU (weak scaling, BASELINE config 4, the default);
cfg5_h{50,100,200} are BASELINE config 5 as written: 1 048 576 instances in total, cut into N
contiguous shards (strong scaling).

--impl reference times the UNMODIFIED reference PyTorch implementation (staged under git-ignored
oracle/_ref by oracle/fetch_ref.py) on the host cores: the batched composition oracle of SURVEY.md
8(c), a bounded sample of the same workload per step, at torch.set_num_threads(1) and
=os.cpu_count(), reporting the better.  Only if oracle/_ref is absent does it fall back to the C
port of the oracle (labelled kind "port").
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

CFG5_TOTAL = 1 << 20
# BASELINE.json configs; FLOP model from SURVEY.md section 8(d) (1 FMA = 2 FLOP, no recompute)
WORKLOADS = {
    # name: (golden fixture with the weights, kind, instances, H, iters, integrator, lr, scaling, description)
    "cfg4_rk4": ("cartpole_h256", "phnn", 65536, 50, 20, "rk4", 0.015, "weak",
                 "BASELINE cfg4: 65536 cart-pole MPC instances x H=50 x 20 Adam iters, RK4, pHNN hidden 256 "
                 "(random init seed 1234)"),
    "cfg4_euler": ("cartpole_h256", "phnn", 65536, 50, 20, "euler", 0.015, "weak",
                   "BASELINE cfg4 (Euler variant): 65536 instances x H=50 x 20 iters, pHNN hidden 256"),
    "cfg3": ("canonical", "canonical", 1024, 10, 50, "euler", 0.03, "weak",
             "BASELINE cfg3: canonical pHNN pole-stabilisation MPC, 1024 instances x H=10 x 50 iters, Euler"),
    "cfg5_h50": ("cartpole_h256", "phnn", CFG5_TOTAL, 50, 20, "rk4", 0.015, "strong",
                 "BASELINE cfg5: 1 048 576 cart-pole MPC instances in total x H=50 x 20 iters, RK4, hidden 256, sharded"),
    "cfg5_h100": ("cartpole_h256", "phnn", CFG5_TOTAL, 100, 20, "rk4", 0.015, "strong",
                  "BASELINE cfg5: 1 048 576 cart-pole MPC instances in total x H=100 x 20 iters, RK4, hidden 256, sharded"),
    "cfg5_h200": ("cartpole_h256", "phnn", CFG5_TOTAL, 200, 20, "rk4", 0.015, "strong",
                  "BASELINE cfg5: 1 048 576 cart-pole MPC instances in total x H=200 x 20 iters, RK4, hidden 256, sharded"),
    "small": ("cartpole_h256", "phnn", 9472, 50, 2, "rk4", 0.015, "weak",
              "profiling-sized cfg4 slice (one full wave, 64 instances per SM): 9472 instances x H=50 x 2 iters, "
              "RK4, hidden 256"),
}
L2_NOTE = ("256 MiB buffer written between timed steps (L2 flush); the per-step working set (activation tape, stage "
           "checkpoints) also exceeds L2")


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


def workload_config(name, wl, world):
    fixture, kind, Btot, H, iters, integ, lr, scaling, desc = wl
    per_gpu = Btot // world if scaling == "strong" else Btot
    total = per_gpu * world
    h = int(load_fixture(fixture)["H_net.net.0.weight"].shape[0])
    return {"workload": desc, "name": name, "instances_per_gpu": per_gpu, "instances_total": total, "horizon": H,
            "iters": iters, "integrator": integ, "hidden": h, "model_kind": kind,
            "parallelism": "instance-sharded x%d" % world, "l2": L2_NOTE}, per_gpu, h


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons during the timed region (pynvml, else nothing)"""

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


# ---------------------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------------------
def port_leg(wl, steps, warmup, threads=None, budget_s=20.0):
    """the C port of the oracle (oracle/libphnn_oracle.so) on a bounded sample: a labelled extra, and the stand-in for
    the reference arm only when oracle/_ref is not staged"""
    from oracle.phnn_oracle import OracleModel, set_threads, max_threads
    fixture, kind, B, H, iters, integ, lr, scaling, desc = wl
    sd = load_fixture(fixture)
    M = OracleModel(sd, kind)
    cores = threads or os.cpu_count() or max_threads()
    set_threads(cores)
    c = cost_for(kind)
    C = M.cost_struct(c["Q"], c["R"], np.zeros(4), c["u_min"], c["u_max"])
    x0 = make_inputs(max(cores * 8, 64), kind, 7).numpy()
    nb = cores
    t0 = time.perf_counter()
    M.mpc_solve(C, x0[:nb], np.zeros((nb, H, 1), np.float32), 0.02, integ, lr=lr, iters=1,
                return_mode="last" if kind == "phnn" else "best")
    per_round = (time.perf_counter() - t0) * iters   # seconds for `cores` full solves
    rounds = max(1, int(budget_s / max(per_round, 1e-3) / max(1, steps + warmup)))
    nb = min(x0.shape[0], cores * rounds)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        M.mpc_solve(C, x0[:nb], np.zeros((nb, H, 1), np.float32), 0.02, integ, lr=lr, iters=iters,
                    return_mode="last" if kind == "phnn" else "best")
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    return {"value": nb / (ms / 1e3), "unit": "solves/s", "cores": int(cores), "kind": "port",
            "sample": "%d of the workload's instances per step (full H=%d, %d iters, %s), %d step(s), C port of the "
                      "oracle (scalar -O2, an accuracy oracle, not the reference)" % (nb, H, iters, integ, len(times)),
            "ms_per_step": ms, "B_sample": int(nb), "step_ms": [1e3 * t for t in times]}


def ref_leg(wl, steps, warmup, budget_s, batch=0):
    """the unmodified reference PyTorch path (composition oracle) on the host cores.

    Calibrates one Adam iteration at 1 and at os.cpu_count() torch threads, keeps the better setting, sizes the
    per-step sample so that warmup + steps full solves fit `budget_s`, then times every step (a full
    `iters`-iteration solve of the sample, cold start, the workload's H / integrator / cost)."""
    import torch
    from oracle import ref_torch
    fixture, kind, B, H, iters, integ, lr, scaling, desc = wl
    sd = load_fixture(fixture)
    model = ref_torch.load_model(sd, kind)
    cores = os.cpu_count() or 1
    c = cost_for(kind)
    mode = "last" if kind == "phnn" else "best"
    xall = make_inputs(4096, kind, 7).numpy()
    calib = {}
    ref_torch.time_solve(model, xall[:32], H, 0.02, integ, c, lr, 1, cores, mode)   # first-call overheads
    for thr, nb in ((1, 128), (cores, 1024)):
        t = ref_torch.time_solve(model, xall[:nb], H, 0.02, integ, c, lr, 1, thr, mode)
        calib[thr] = nb / t                                                       # instance-iterations per second
    best_thr = max(calib, key=calib.get)
    rate = calib[best_thr] / iters                                                # full solves per second (estimate)
    n_steps = max(1, steps + warmup)
    nb = batch or int(rate * budget_s / n_steps)
    nb = int(min(4096, max(64, nb // 64 * 64)))
    times = []
    for s in range(warmup + steps):
        t = ref_torch.time_solve(model, xall[:nb], H, 0.02, integ, c, lr, iters, best_thr, mode)
        if s >= warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    spread = (max(times) - min(times)) / np.mean(times) if len(times) > 1 else 0.0
    return {"value": nb / (ms / 1e3), "unit": "solves/s", "cores": int(cores), "torch_threads": int(best_thr),
            "kind": "reference-pytorch", "torch": torch.__version__,
            "sample": "%d of the workload's instances per step (full H=%d, %d Adam iters, %s, cold start), %d timed step(s) "
                      "after %d warm-up; unmodified reference PyTorch (integrators.rollout_trajectory_differentiable + "
                      "backward + torch.optim.Adam) from oracle/_ref" % (nb, H, iters, integ, len(times), warmup),
            "calibration_instance_iters_per_s": {"threads_%d" % k: v for k, v in calib.items()},
            "ms_per_step": ms, "B_sample": nb, "step_ms": [1e3 * t for t in times], "spread": float(spread)}


def reference_available():
    try:
        from oracle import fetch_ref
        return fetch_ref.available()
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg4_rk4", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override instances (per GPU for weak, total for strong workloads)")
    ap.add_argument("--scaling", default="", choices=["", "weak", "strong"], help="override the workload's scaling mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the plain-TF32 (tensor_mode 1) side measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing parity sample")
    ap.add_argument("--quick", action="store_true",
                    help="long workloads (cfg5): no separate kernel-only re-timing (kernel time = step time - exchange), e2e over one step")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer"],
                    help="final exchange at N>1: NCCL all_gather, or the fused peer-store epilogue of the solve kernel")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="seconds of CPU time for the whole reference arm")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm: instances per step (default: sized to --ref-budget)")
    ap.add_argument("--tensor-pair", type=int, default=-1,
                    help="1: solve jobs run as CTA pairs (tcgen05 cta_group::2, same results); default: the library's (0)")
    ap.add_argument("--tensor-mode", type=int, default=-1,
                    help="0 FP32-FMA kernel, 4 tcgen05 3 x FP16 hi/lo products with A in TMEM (default), 2 tcgen05 TF32 + BF16 "
                         "correction product, 3 tcgen05 3xTF32, 1 tcgen05 plain TF32")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.batch:
        wl[2] = args.batch
    if args.scaling:
        wl[7] = args.scaling
    fixture, kind, Btot, H, iters, integ, lr, scaling, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    S = 4 if integ == "rk4" else 1
    sd = load_fixture(fixture)
    config, B, h = workload_config(args.workload, wl, world)

    if args.impl == "reference":
        if rank != 0:
            return
        if reference_available():
            res = ref_leg(wl, args.steps, args.warmup, args.ref_budget, args.ref_batch)
        else:
            res = port_leg(wl, args.steps, min(args.warmup, 1), budget_s=60.0)
            res["note"] = "oracle/_ref not staged: C port of the oracle timed instead of the reference"
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample") if k in res}
        for k in ("torch_threads", "torch", "calibration_instance_iters_per_s", "spread", "note"):
            if k in res:
                cpu[k] = res[k]
        line = {"impl": "reference", "metric": "cartpole_mpc_solves_per_s", "value": res["value"], "unit": "solves/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cpu,
                "e2e": {"value": res["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "step_ms": res.get("step_ms")}
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
    if args.tensor_pair >= 0:
        pk.set_option("tensor_pair", args.tensor_pair)
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

    # ---- the final exchange: fused peer-store epilogue where it can be set up, else NCCL all_gather ----
    peer = None
    gather_kind = "none"
    if world > 1:
        gather_kind = "nccl-all_gather"
        if args.gather in ("auto", "peer"):
            try:
                from phnn_mpc_b200.peer import PeerGather
                peer = PeerGather(B, H, world, rank, dev)
                gather_kind = "peer-store epilogue (solve kernel writes U*/cost into every rank's result buffer over NVLink)"
            except Exception as ex:   # noqa: BLE001 -- reported in the line
                if args.gather == "peer":
                    raise
                gather_kind = "nccl-all_gather (peer-store setup failed: %s)" % (repr(ex)[:120],)
                peer = None

    def step_device():
        if peer is not None:
            o = mpc.solve(x0, peer=peer)
            peer.finish()
            return o, peer.U_all
        o = mpc.solve(x0)
        Uall = gather_shards(o["U"], B * world) if world > 1 else o["U"]
        return o, Uall

    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed between ----
    for _ in range(args.warmup):
        out, Uall = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, Uall = step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    assert Uall.shape[0] == B * world
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t = torch.tensor([float(sum(step_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = B * world / (ms_per_step / 1e3)

    # the exchange alone, warmed, median of 7 (it is issued once per step and already inside the step time above)
    gather_ms = 0.0
    if world > 1:
        ts = []
        for i in range(10):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if peer is not None:
                peer.exchange_only(out["U"], out["best_cost"])
            else:
                gather_shards(out["U"], B * world)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        gather_ms = float(np.median(ts))

    # kernel-only time of the solve (no exchange) for the roofline: same launches, events around the solve alone
    kts = []
    for _ in range(0 if args.quick else min(args.steps, 3)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mpc.solve(x0)
        e1.record()
        torch.cuda.synchronize()
        kts.append(e0.elapsed_time(e1))
    kernel_ms = float(np.mean(kts)) if kts else float(np.mean(step_ms)) - gather_ms

    # ---- end-to-end through the public API with host buffers: every step, L2 flushed between steps ----
    e2e = None
    if not args.no_e2e:
        U_host = torch.empty((B * world if world > 1 else B, H, 1), dtype=torch.float32).pin_memory()
        c_host = torch.empty((B,), dtype=torch.float32).pin_memory()
        tot = 0.0
        e2e_steps = 1 if args.quick else args.steps
        # one untimed step of this exact path first: the first pinned-host -> device copy of a process takes ~60 ms
        # (measured, tools/gpu_e2e_diag.py), which is set-up, not the step
        for i in range(e2e_steps + 1):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            xd = x0_host.to(dev, non_blocking=True)
            if peer is not None:
                o = mpc.solve(xd, peer=peer)
                peer.finish()
                Ue = peer.U_all
            else:
                o = mpc.solve(xd)
                Ue = gather_shards(o["U"], B * world) if world > 1 else o["U"]
            U_host.copy_(Ue, non_blocking=True)
            c_host.copy_(o["best_cost"], non_blocking=True)
            torch.cuda.synchronize()
            if i > 0:
                tot += time.perf_counter() - t0
        barrier()
        t = torch.tensor([1e3 * tot / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world / (float(t.item()) / 1e3), "unit": "solves/s",
               "h2d_bytes_per_step": int(x0_host.numel() * 4), "d2h_bytes_per_step": int((U_host.numel() + B) * 4),
               "steps": e2e_steps, "ms_per_step": float(t.item()),
               "note": "pinned x0 -> device, solve (+ exchange), U of the whole job and this rank's costs -> pinned host, "
                       "wall clock per step, max over ranks, L2 flushed between steps, after one untimed step of the same path"}

    # ---- side measurement: one FP16 product per algorithmic product (tensor_mode 5), the looser stated-tolerance path
    #      north_star permits for a reduced-precision tensor-core path; NOT the headline ----
    alt = None
    tmode = pk.get_option("tensor_mode")
    uses_tc = tmode in (1, 2, 3, 4, 5) and B >= pk.get_option("tensor_min_batch")
    if uses_tc and tmode != 5 and not args.no_alt:
        pk.set_option("tensor_mode", 5)
        for _ in range(2):
            mpc.solve(x0)
        ats = []
        for _ in range(min(args.steps, 3)):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o1 = mpc.solve(x0)
            e1.record()
            torch.cuda.synchronize()
            ats.append(e0.elapsed_time(e1))
        pk.set_option("tensor_mode", tmode)
        a_ms = float(np.mean(ats))
        t = torch.tensor([a_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        alt = {"label": "tensor_mode=5: second-generation tcgen05 kernel with ONE FP16 product per algorithmic product "
                        "(operands rounded to FP16, a third of the tensor work; NOT the headline; looser stated tolerance "
                        "2e-4 cost / 1e-3 dJ/dU, north_star allows a stated looser bound for a reduced-precision path)",
               "ms_per_step": float(t.item()),
               "value": B * world / (float(t.item()) / 1e3), "unit": "solves/s (solve kernel only, no exchange)",
               "U_maxabs_vs_headline_mode": float((o1["U"] - out["U"]).abs().max().item()),
               "best_cost_rel_vs_headline_mode": float(((o1["best_cost"] - out["best_cost"]).abs().max() /
                                                       out["best_cost"].abs().max()).item())}

    # ---- parity sample: the first 64 instances of this rank's timed output vs the oracle and the reference ----
    parity = None
    if not args.no_parity and rank == 0:
        parity = parity_sample(mpc, x0, out, wl, B)

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
    achieved = algo / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
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
        tf32_flops = base * (3 if tmode == 3 else 1) if tmode < 4 else 0.0
        # 16-bit MMAs (twice the TF32 rate): mode 2 one BF16 product of twice the depth; mode 4 three FP16 products
        # (a_hi b_hi + a_hi b_lo + a_lo b_hi), operand A read from tensor memory
        bf16_flops = base * 2 if tmode == 2 else (base * 3 if tmode == 4 else (base if tmode == 5 else 0.0))
        mma_flops = tf32_flops + bf16_flops
        tf32_equiv_flops = tf32_flops + bf16_flops / 2
        scheme = {5: "1 x FP16 product, A in TMEM (looser stated tolerance)", 4: "3 x FP16 hi/lo products, A in TMEM (FP32-level accuracy)", 3: "3xTF32 error-compensated",
                  2: "TF32 + BF16 correction product (FP32-level accuracy)", 1: "plain TF32"}[tmode]
        # HBM side of the same kernel: the activation tape (a1, a2, g1: 3 h floats per instance and evaluation) is
        # written by the forward sweep and read back by the adjoint
        tape_bytes = tiles * iters * H * S * 2.0 * 3 * h * 128 * 4
        hbm_peak = peaks.get("hbm_gbs", 6500.0) if peaks else 6500.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": achieved / bf16_peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, this pool)" if peaks else
                                   "fallback 1400 (B200_PROFILING.md)",
                    "algorithmic_flops_per_launch": algo,
                    "kernel": ("phnn_tc16_kernel<MK,NS,HID> (tcgen05 kind::f16, %s), one launch per step" % scheme) if tmode >= 4 else
                              "phnn_tc_kernel<MK,NS,HID> (tcgen05 kind::tf32%s, %s), one launch per step" % (
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
                    "note": ("FP32-level accuracy on the tensor cores costs three FP16 products per algorithmic product (hi/lo "
                             "split operands): frac vs the bf16 peak is bounded by 1/3; " if tmode == 4 else
                             "one FP16 product per algorithmic product (tensor_mode 5, looser stated tolerance); " if tmode == 5 else
                             "FP32-level accuracy on the tensor cores costs one TF32 product (half the bf16 rate) plus a BF16 "
                             "correction product of twice the depth: frac vs the bf16 peak is bounded by 1/4; ") +
                            "the adjoint reads the forward activations from an HBM tape instead of recomputing them (4 tensor "
                            "products per pair instead of 6)"}
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
        r_tflops = Bp * Tp * 73360 / (rms * 1e-3) / 1e12
        # the same job on the latency kernel (the route of small batches), for the record
        fwd_min = pkp.get_option("tensor_fwd_min_batch")
        lat_ms = None
        if fwd_min > 0:
            pkp.set_option("tensor_fwd_min_batch", 0)
            for _ in range(2):
                _rollout(pkp, xp, Up, 0.05, "rk4")
            torch.cuda.synchronize()
            tl = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _rollout(pkp, xp, Up, 0.05, "rk4")
                e1.record()
                torch.cuda.synchronize()
                tl.append(e0.elapsed_time(e1))
            lat_ms = float(np.median(tl))
            pkp.set_option("tensor_fwd_min_batch", fwd_min)
        on_tc = fwd_min > 0 and Bp >= fwd_min
        # parity sample of the timed route: first 64 instances, first 10 steps, against the CPU oracle
        r_par = None
        if not args.no_parity:
            from oracle.phnn_oracle import OracleModel
            tr_g = _rollout(pkp, xp, Up, 0.05, "rk4")
            tr_g = (tr_g[0] if isinstance(tr_g, (tuple, list)) else tr_g)[:64, :11].cpu().numpy()
            tr_o = OracleModel(sdp, "phnn").rollout(xp[:64].cpu().numpy(), Up[:64, :10].cpu().numpy(), 0.05, "rk4")
            r_par = {"n": 64, "steps": 10, "traj_rel": float(np.abs(tr_g - tr_o).max() / np.abs(tr_o).max()), "tol": 1e-4}
        rollout_metric = {"metric": "phnn_rk4_rollout_instance_steps_per_s", "value": Bp * Tp / (rms * 1e-3), "unit": "instance-steps/s",
                          "ms": rms, "config": "BASELINE cfg2: pendulum pHNN (shipped weights, h=64, learned G), 4096 initial states x H=100, "
                                               "RK4, dt 0.05, one launch (%s)" % (
                                                   "phnn_tc16_kernel<PHNN_GNET,2,64>: forward-only tcgen05 instantiation, 3 x FP16 hi/lo products, "
                                                   + ("64 tiles of 64 instances on 64 SMs" if pkp.get_option("tensor_fwd_sparse") == 1 and Bp <= 64 * 148
                                                      else "32 tiles of 128 instances on 32 SMs") if on_tc else "latency kernel, FP32 FMA"),
                          "latency_kernel_ms": lat_ms, "parity_sample": r_par,
                          "roofline": {"bound": "latency (400 sequential evaluations per tile)", "achieved": r_tflops, "peak": fp32_peak,
                                       "unit": "TFLOP/s", "frac": r_tflops / fp32_peak, "algorithmic_flops_per_launch": Bp * Tp * 73360,
                                       "note": "one CTA per tile runs 100 steps x 4 stages one after the other (%.1f us per evaluation "
                                               "of a tile); the same launch takes the same time up to one tile per SM (9472 instances on "
                                               "64-instance tiles, 18944 on 128-instance tiles in 1.9 ms). frac = algorithmic FLOP/s over "
                                               "the FP32-FMA rate measured in this run" % (rms * 1e3 / (Tp * 4))}}
    except Exception as ex:  # the headline line must still be printed
        rollout_metric = {"error": repr(ex)}

    cpu = None
    if not args.no_cpu_baseline:
        if reference_available():
            r = ref_leg(wl, 1, 0, 20.0)
        else:
            r = port_leg(wl, 1, 0, budget_s=15.0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "torch_threads",
                                 "calibration_instance_iters_per_s") if k in r}

    line = {"metric": "cartpole_mpc_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "kernel_path": {5: "tcgen05-1xFP16-A-in-TMEM", 4: "tcgen05-3xFP16-A-in-TMEM", 3: "tcgen05-3xTF32", 2: "tcgen05-TF32+BF16corr", 1: "tcgen05-TF32"}[tmode] if uses_tc else "fp32-fma",
            "exchange": {"kind": gather_kind, "ms": gather_ms,
                         "note": "issued once per step, inside every timed step; `ms` is the exchange alone, warmed, median of 7"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "tensor_pair": int(pk.get_option("tensor_pair")),
            "roofline": roofline, "cpu_baseline": cpu, "parity_sample": parity, "fp16_mode5": alt,
            "rollout": rollout_metric, "step_ms": step_ms}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_sample(mpc, x0, out, wl, B, n=64):
    """After timing: the first `n` instances of the timed step's output against (a) the CPU oracle run on the same
    instances and (b), for cfg4_rk4 on rank 0, the fixture recorded from the reference itself
    (tests/golden/cfg4_shape.npz: same weights, same first 64 instances, H=50, RK4, 20 iterations).  The cost
    history comes from one more (untimed) solve of the whole batch with the history requested."""
    import torch
    from oracle.phnn_oracle import OracleModel, set_threads
    fixture, kind, Btot, H, iters, integ, lr, scaling, desc = wl
    n = min(n, B)
    o2 = mpc.solve(x0, want_hist=True)
    torch.cuda.synchronize()
    same = bool(torch.equal(o2["U"], out["U"]))
    U = out["U"][:n].cpu().numpy()
    hist = o2["cost_hist"][:, :n].cpu().numpy()
    res = {"n": int(n), "rerun_bit_identical": same}
    try:
        sd = load_fixture(fixture)
        M = OracleModel(sd, kind)
        set_threads(os.cpu_count() or 1)
        c = cost_for(kind)
        C = M.cost_struct(c["Q"], c["R"], np.zeros(4), c["u_min"], c["u_max"])
        mode = "last" if kind == "phnn" else "best"
        Uo, histo, _ = M.mpc_solve(C, x0[:n].cpu().numpy(), np.zeros((n, H, 1), np.float32), 0.02, integ, lr=lr,
                                   iters=iters, return_mode=mode)
        res["vs_oracle"] = {"hist_rel": float(np.abs(hist - histo).max() / np.abs(histo).max()),
                            "U_abs": float(np.abs(U - Uo).max()), "U_tol": 0.02 * lr + 1e-5, "hist_tol": 1e-4}
    except Exception as ex:  # noqa: BLE001
        res["vs_oracle"] = {"error": repr(ex)}
    gpath = os.path.join(REPO, "tests", "golden", "cfg4_shape.npz")
    if fixture == "cartpole_h256" and integ == "rk4" and os.path.exists(gpath):
        z = np.load(gpath)
        if int(z["H"]) == H and int(z["iters"]) == iters and np.array_equal(z["x0"][:n], x0[:n].cpu().numpy()):
            res["vs_reference_golden"] = {"hist_rel": float(np.abs(hist - z["rk4_hist"][:, :n]).max() / np.abs(z["rk4_hist"]).max()),
                                          "U_abs": float(np.abs(U - z["rk4_U_last"][:n]).max()),
                                          "source": "tests/golden/cfg4_shape.npz (reference PyTorch, recorded by make_golden.py)"}
    return res


if __name__ == "__main__":
    main()
