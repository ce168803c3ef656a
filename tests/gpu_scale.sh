#!/bin/bash
# usage: gpu_scale.sh <ngpus> <workload> <steps> <warmup> [test]   (one bench line at N GPUs; "test" = multi-GPU pytest first)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=$1; W=$2; S=$3; WU=$4
if [ "$5" == "test" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_$N.log
  tail -4 gpurun_out/pytest_multi_$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload $W --steps $S --warmup $WU --no-cpu > gpurun_out/scale_${W}_$N.json 2> gpurun_out/scale_${W}_$N.err
echo "bench exit $?"; grep -v "^\[bench\]\|Makefile" gpurun_out/scale_${W}_$N.err | tail -3
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${W}_$N.json").read().strip().split("\n")[-1])
    print("$W x$N", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,2) for k,v in d["phase_ms_per_step"].items()}, {k:round(v,2) for k,v in d["sweep_detail_ms_per_step"].items()}, "e2e", d["e2e"] and d["e2e"]["breakdown"])
except Exception as e: print("parse failed", e)
PY
