# 8-GPU runs of the round: the driver's own command for cfg4 (weak scaling, fused peer-store exchange by default), the same
# with the plain NCCL all_gather, the reference arm under torchrun, and cfg5 as written (strong scaling, 1 048 576 instances)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%s: value %.0f %s ms/step %.1f e2e %.0f exchange %s clocks %s" % (sys.argv[2], d["value"], d.get("unit"), d["ms_per_step"], d["e2e"]["value"],
          json.dumps(d.get("exchange"))[:150], json.dumps(d.get("clocks"))))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
}
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/s8_cfg4.err | grep '^{' > gpurun_out/s8_cfg4_n$N.json; show gpurun_out/s8_cfg4_n$N.json "cfg4 weak N=$N (default exchange)"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-alt --gather nccl 2> gpurun_out/s8_cfg4_nccl.err | grep '^{' > gpurun_out/s8_cfg4_nccl_n$N.json; show gpurun_out/s8_cfg4_nccl_n$N.json "cfg4 weak N=$N (NCCL all_gather)"
for w in cfg5_h50 cfg5_h100 cfg5_h200; do
  $TR bench.py --gpus $N --workload $w --steps 2 --warmup 1 --quick --no-alt --no-parity --no-cpu-baseline 2> gpurun_out/s8_${w}.err | grep '^{' > gpurun_out/s8_${w}_n$N.json
  show gpurun_out/s8_${w}_n$N.json "$w strong N=$N"
done
