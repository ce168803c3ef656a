#!/bin/bash
# usage: gpu_multi2.sh <ngpus> <workload> <steps> <warmup>   (multi-GPU parity test, then one bench line)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=$1; W=$2; S=$3; WU=$4
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --workload $W --steps $S --warmup $WU --no-cpu --no-e2e > gpurun_out/scale_${W}_$N.json 2> gpurun_out/scale_${W}_$N.err
echo "bench exit $?"; tail -3 gpurun_out/scale_${W}_$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${W}_$N.json").read().strip().split("\n")[-1])
    print("$W x$N", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,3) for k,v in d["phase_ms_per_step"].items()}, d.get("sweep_detail_ms_per_step"))
except Exception as e: print("parse failed", e)
PY
