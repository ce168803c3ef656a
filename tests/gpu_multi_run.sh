#!/bin/bash
# usage: gpu_multi_run.sh <ngpus> <workload> <steps> <warmup> [extra]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=$1; W=$2; S=$3; WU=$4; shift 4
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -8 gpurun_out/pytest_multi.log
for n in $N; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --workload $W --steps $S --warmup $WU "$@" > gpurun_out/scale_${W}_1.json 2> gpurun_out/scale_${W}_1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --workload $W --steps $S --warmup $WU "$@" > gpurun_out/scale_${W}_$n.json 2> gpurun_out/scale_${W}_$n.err
    EALS_PEER_STORE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $n --workload $W --steps $S --warmup $WU "$@" > gpurun_out/scale_${W}_${n}_nccl.json 2> gpurun_out/scale_${W}_${n}_nccl.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${W}_${n}_nccl.json").read().strip().split("\n")[-1])
    print("$W x$n NCCL-broadcast exchange", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"])
except Exception as e: print("parse failed", e)
PY
  fi
  echo "bench $W x$n exit $?"; tail -3 gpurun_out/scale_${W}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${W}_$n.json").read().strip().split("\n")[-1])
    print("$W x$n", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["phase_ms_per_step"].items()}, d.get("sweep_detail_ms_per_step"))
except Exception as e: print("parse failed", e)
PY
done
