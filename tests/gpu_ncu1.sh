#!/bin/bash
# one full capture: gpu_ncu1.sh <workload> <tag> <kernel regex> <skip> [count]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=$1; TAG=$2; KR=$3; SK=$4; CNT=${5:-1}
CMD="python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KR -s $SK -c $CNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1; echo "capture exit $?"
ls -la gpurun_out/prof_$TAG.ncu-rep
