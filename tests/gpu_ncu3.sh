#!/bin/bash
# three full captures on one workload: warp kernel, one-CTA kernel, heavy step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=${1:-c4}; TAG=${2:-r1b}
CMD="python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:cd_warp_kernel -s ${3:-8} -c 1 -o gpurun_out/prof_${TAG}_warp -f $CMD > gpurun_out/ncu_${TAG}_warp.log 2>&1; echo "warp capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:cd_row_block -s ${4:-7} -c 1 -o gpurun_out/prof_${TAG}_mid -f $CMD > gpurun_out/ncu_${TAG}_mid.log 2>&1; echo "mid capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:heavy_step -s ${5:-10000} -c 2 -o gpurun_out/prof_${TAG}_heavy -f $CMD > gpurun_out/ncu_${TAG}_heavy.log 2>&1; echo "heavy capture exit $?"
ls -la gpurun_out/*.ncu-rep
