#!/bin/bash
# usage: gpu_env_sweep.sh <workload> <steps> <warmup> VAR v1 v2 ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=$1; S=$2; WU=$3; VAR=$4; shift 4
for v in "$@"; do
  env $VAR=$v timeout 600 python bench.py --workload $W --steps $S --warmup $WU --no-e2e --no-cpu > gpurun_out/sweep_${VAR}_$v.json 2> gpurun_out/sweep_${VAR}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_${VAR}_$v.json").read().strip().split("\n")[-1])
    print("$VAR=$v", "ms/step %.2f"%d["ms_per_step"], "launches", d["gpu_launches"], {k:round(x,1) for k,x in d["sweep_detail_ms_per_step"].items()})
except Exception as e: print("$VAR=$v parse failed", e)
PY
done
