"""Measures every BASELINE.json config on one GPU (device-resident timing with CUDA events) next to the
CPU oracle port on the host cores; writes JSON lines (one per config) to stdout / gpurun_out/."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench as B
from phnn_mpc_b200.batched import BatchedMPC, CostSpec, rollout
from phnn_mpc_b200.packing import PackedModel
from oracle.phnn_oracle import OracleModel, set_threads

dev = torch.device("cuda", 0)
cores = os.cpu_count()
set_threads(cores)


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = []
# cfg2: pendulum RK4 rollout 4096 x 100
sd = B.load_fixture("pendulum")
pk = PackedModel({k: torch.from_numpy(v) for k, v in sd.items()}, "phnn", device=dev)
g = torch.Generator().manual_seed(1)
Bn, T = 4096, 100
x0 = torch.stack([(torch.rand(Bn, generator=g) * 2 - 1) * np.pi, torch.rand(Bn, generator=g) * 2 - 1], 1).to(dev)
U = (torch.rand(Bn, T, 1, generator=g) * 4 - 2).to(dev)
ms = timed(lambda: rollout(pk, x0, U, 0.05, "rk4"))
M = OracleModel(sd, "phnn")
t0 = time.perf_counter(); M.rollout(x0.cpu().numpy(), U.cpu().numpy(), 0.05, "rk4"); cpu_s = time.perf_counter() - t0
out.append({"config": "cfg2 pendulum RK4 rollout 4096 x H=100 (h=64, learned G)", "ms": ms, "inst_steps_per_s": Bn * T / ms * 1e3,
            "flops_algo": Bn * T * 73360, "tflops": Bn * T * 73360 / ms / 1e9, "cpu_port_inst_steps_per_s": Bn * T / cpu_s, "cpu_cores": cores})

# cfg1: single-instance controller latency (drop-in MPCController) and cfg3 B=1
sys.path.insert(0, os.path.join(R, "tests"))
from phnn_mpc_b200.dropin.pHNN import pHNN
from phnn_mpc_b200.dropin.pHNN_canonical import pHNN_Canonical
from phnn_mpc_b200.dropin.mpc_controller import MPCController
from phnn_mpc_b200.dropin.mpc_controller_canonical import create_mpc_controller
import yaml
torch.manual_seed(0)
m = pHNN(os.path.join(R, "configs", "cartpole_phnn.yaml"))
c = MPCController(m, 20, 0.02, [10.0, 200.0, 1.0, 10.0], 0.01, [0, 0, 0, 0], -15.0, 15.0, lr=0.015, max_iterations=30)
s = np.array([0.0, 0.1, 0.0, 0.0])
c.compute_control(s)
t0 = time.perf_counter()
for _ in range(5): c.compute_control(s)
ms1 = (time.perf_counter() - t0) / 5 * 1e3
sdm = {k: v.detach().numpy() for k, v in m.state_dict().items()}
Mo = OracleModel(sdm, "phnn"); set_threads(1)
C = Mo.cost_struct([10.0, 200.0, 1.0, 10.0], [0.01], np.zeros(4), -15.0, 15.0)
t0 = time.perf_counter(); Mo.mpc_solve(C, s[None].astype(np.float32), np.zeros((1, 20, 1), np.float32), 0.02, "euler", lr=0.015, iters=30); cpu1 = time.perf_counter() - t0
set_threads(cores)
out.append({"config": "cfg1 cart-pole pHNN MPC single instance (H=20, 30 it, h=128) via drop-in MPCController.compute_control (host in/out)",
            "ms_per_solve": ms1, "solves_per_s": 1e3 / ms1, "cpu_port_1thread_ms": cpu1 * 1e3})

# cfg3: canonical, 1024 instances
sd3 = B.load_fixture("canonical")
pk3 = PackedModel({k: torch.from_numpy(v) for k, v in sd3.items()}, "canonical", device=dev)
x3 = B.make_inputs(1024, "canonical", 3).to(dev)
mpc3 = BatchedMPC(pk3, 10, 0.02, CostSpec.make(4, 1, [0.0, 1000.0, 0.0, 100.0], [1e-4], None, -30.0, 30.0), integrator="euler", lr=0.03, iters=50, return_mode="best")
ms3 = timed(lambda: mpc3.solve(x3))
M3 = OracleModel(sd3, "canonical")
C3 = M3.cost_struct([0.0, 1000.0, 0.0, 100.0], [1e-4], np.zeros(4), -30.0, 30.0)
t0 = time.perf_counter(); M3.mpc_solve(C3, x3.cpu().numpy(), np.zeros((1024, 10, 1), np.float32), 0.02, "euler", lr=0.03, iters=50, return_mode="best"); cpu3 = time.perf_counter() - t0
out.append({"config": "cfg3 canonical pole-stabilisation MPC, 1024 instances (H=10, 50 it, Euler, h=128)", "ms": ms3, "solves_per_s": 1024 / ms3 * 1e3,
            "tflops": 1024 * 6.78e7 / ms3 / 1e9, "cpu_port_solves_per_s": 1024 / cpu3, "cpu_cores": cores})

# cfg4 / cfg5 shards (h=256)
sd4 = B.load_fixture("cartpole_h256")
pk4 = PackedModel({k: torch.from_numpy(v) for k, v in sd4.items()}, "phnn", device=dev)
spec = CostSpec.make(4, 1, [10.0, 200.0, 1.0, 10.0], [0.01], None, -15.0, 15.0)
for name, Bn, H, integ in (("cfg4 Euler 65536 x H=50 x 20 it", 65536, 50, "euler"), ("cfg4 RK4 65536 x H=50 x 20 it", 65536, 50, "rk4"),
                           ("cfg5 shard (1M/8) RK4 131072 x H=50", 131072, 50, "rk4"), ("cfg5 shard (1M/8) RK4 131072 x H=100", 131072, 100, "rk4"),
                           ("cfg5 shard (1M/8) RK4 131072 x H=200", 131072, 200, "rk4")):
    x4 = B.make_inputs(Bn, "phnn", 7).to(dev)
    mpc = BatchedMPC(pk4, H, 0.02, spec, integrator=integ, lr=0.015, iters=20)
    ms4 = timed(lambda: mpc.solve(x4), reps=2, warm=1)
    S = 4 if integ == "rk4" else 1
    fl = B.solve_flops("phnn", 256, 4, H, 20, S) * Bn
    out.append({"config": name + " (h=256, default tcgen05 route)", "ms": ms4, "solves_per_s": Bn / ms4 * 1e3, "algorithmic_tflops": fl / ms4 / 1e9})
    del x4
for o in out:
    print(json.dumps(o))
