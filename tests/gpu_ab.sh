#!/bin/bash
# usage: gpu_ab.sh <workload> "<tag>:<ENV=V ENV2=V2>;<tag2>:..."  [ncu-spec ...]
#   runs the GPU parity tests, then one short bench per env configuration (A/B comparison of
#   kernel variants), then optional `ncu --set full` captures: each ncu-spec = tag,kernel-regex,skip[,ENV=V...]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=$1; CFGS=$2; shift 2
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
grep -q "pytest exit 0" gpurun_out/pytest_gpu.log || { grep -B30 "Error\|assert" gpurun_out/pytest_gpu.log | tail -60; exit 1; }
IFS=';' read -ra RUNS <<< "$CFGS"
for r in "${RUNS[@]}"; do
  TAG=${r%%:*}; ENVS=${r#*:}
  env $ENVS timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_${W}_$TAG.json 2> gpurun_out/ab_${W}_$TAG.err
  echo "bench $TAG ($ENVS) exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${W}_$TAG.json").read().strip().split("\n")[-1])
    print("  $TAG ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["phase_ms_per_step"].items()}, {k:round(v,1) for k,v in d["sweep_detail_ms_per_step"].items()}, "loss %.10g"%d["loss_after"])
except Exception as e: print("parse failed", e)
PY
done
for spec in "$@"; do
  IFS=',' read -ra P <<< "$spec"
  TAG=${P[0]}; KRE=${P[1]}; SKIP=${P[2]}; ENVS="${P[@]:3}"
  env $ENVS ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o gpurun_out/prof_$TAG -f \
     python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu $TAG exit $?"
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -5
