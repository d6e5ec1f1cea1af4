run() { # name envvars... -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --no-prepass --steps 5 --warmup 3 $ARGS > gpurun_out/$name.json 2> gpurun_out/$name.err || tail -5 gpurun_out/$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]/1e6,1),"M/s", round(d["ms_per_step"],2),"ms", "e2e", round(d["e2e"]["value"]/1e6,1), d.get("kernel_ms_per_rollout"))
except Exception as e: print("$name failed", e)
PY
}
