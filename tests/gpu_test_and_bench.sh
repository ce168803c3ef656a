#!/bin/bash
# usage: gpu_test_and_bench.sh "<workload steps warmup>;<...>"   (runs GPU tests first)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if grep -q "pytest exit 0" gpurun_out/pytest_gpu.log; then
  IFS=';' read -ra RUNS <<< "$1"
  for r in "${RUNS[@]}"; do
    set -- $r
    W=$1; S=$2; WU=$3; shift 3
    timeout 900 python bench.py --workload $W --steps $S --warmup $WU "$@" > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
    echo "bench $W exit $?"; tail -2 gpurun_out/bench_$W.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$W.json").read().strip().split("\n")[-1])
    print("$W", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["phase_ms_per_step"].items()}, "launches", d["gpu_launches"], "e2e", d["e2e"] and "%.3e"%d["e2e"]["value"])
except Exception as e: print("parse failed", e)
PY
  done
fi
