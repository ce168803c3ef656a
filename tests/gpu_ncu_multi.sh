#!/bin/bash
# usage: gpu_ncu_multi.sh <workload> spec...   spec = tag,kernel-regex,skip[,ENV=V...]  (one --set full capture each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=$1; shift
for spec in "$@"; do
  IFS=',' read -ra P <<< "$spec"
  TAG=${P[0]}; KRE=${P[1]}; SKIP=${P[2]}; ENVS="${P[@]:3}"
  env $ENVS ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -o gpurun_out/prof_$TAG -f \
     python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
  echo "ncu $TAG exit $?"
done
ls -la gpurun_out/*.ncu-rep | tail -6
