#!/bin/bash
# usage: gpu_bench_run.sh <workload> <steps> <warmup> [extra bench args]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
W=${1:-c4}; S=${2:-3}; WU=${3:-3}; shift 3
timeout 1200 python bench.py --workload $W --steps $S --warmup $WU "$@" > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
echo "bench $W exit $?"; tail -3 gpurun_out/bench_$W.err; cat gpurun_out/bench_$W.json
