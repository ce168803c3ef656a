#!/bin/bash
# usage: gpu_determinism.sh <ngpus> <runs> [workload]   — the open issue of round 1 (profiles/README.md r01j):
# run the N-GPU bench several times and compare loss_after (after warmup + steps epochs) with each other and,
# for c4 with 3 + 5 epochs, with the single-GPU value 35018277.5368.  Also once with EALS_PEER_PRED_CACHE=0 and
# once with EALS_PEER_STORE=0 to tell the peer stores from the routed caches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=$1; R=${2:-4}; W=${3:-c4}
run() {  # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --workload $W --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$TAG', 'ms/step %.2f' % d['ms_per_step'], 'loss_after', repr(d['loss_after']))"
}
for i in $(seq 1 $R); do run "default#$i" EALS_X=0; done
run "no-peer-pred-cache" EALS_PEER_PRED_CACHE=0
run "nccl-broadcast" EALS_PEER_STORE=0
