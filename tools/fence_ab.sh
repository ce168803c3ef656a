# A/B of the per-thread system fence after peer stores (EALS_PEER_FENCE), 4 GPUs, c4 workload.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S="--steps 8 --warmup 3 --no-cpu --no-configs --no-e2e"
EALS_PEER_FENCE=0 timeout 150 $TR --nproc-per-node 4 --master-port 29811 bench.py --gpus 4 $S > gpurun_out/r2w_4gpu_nofence.json 2> gpurun_out/r2w_4gpu_nofence.err; echo rc=$?
timeout 150 $TR --nproc-per-node 4 --master-port 29812 bench.py --gpus 4 $S > gpurun_out/r2w_4gpu_fence.json 2> gpurun_out/r2w_4gpu_fence.err; echo rc=$?
for f in gpurun_out/r2w_4gpu_nofence.json gpurun_out/r2w_4gpu_fence.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['ms_per_step'], d.get('loss_after'), d.get('replicas_consistent'), d.get('per_rank_phase_ms_per_step'))
PY
done
