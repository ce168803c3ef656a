#!/bin/bash
# usage: gpu_ncu_run.sh <workload> <kernel-regex-for-full-capture> <tag>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=${1:-c3}; KR=${2:-cd_cta}; TAG=${3:-r1}
CMD="python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cd_|gram_|loss_|sumsq|eval_' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KR -s 1 -c 2 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out/
