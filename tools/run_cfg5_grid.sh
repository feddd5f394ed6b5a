N=$1
run() { if [ "$N" = "1" ]; then python bench.py "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N "$@"; fi; }
for w in cfg5_h50 cfg5_h100 cfg5_h200; do
  run --workload $w --steps 2 --warmup 1 --quick --no-alt --no-parity --no-cpu-baseline 2> gpurun_out/r2_${w}_n$N.err | grep '^{' > gpurun_out/r2_${w}_n$N.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_${w}_n$N.json").read().strip().splitlines()[-1])
    print("$w N=$N value %.0f solves/s ms/step %.1f e2e %.0f exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d["exchange"])[:160]))
except Exception as ex:
    print("$w N=$N FAILED", ex)
PY
done
run --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_cfg4_n$N.err | grep '^{' > gpurun_out/r2_cfg4_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_cfg4_n$N.json").read().strip().splitlines()[-1])
print("cfg4 N=$N value %.0f solves/s ms/step %.1f e2e %.0f exchange %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d["exchange"])[:200], json.dumps(d["clocks"])))
PY
