TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
S="--steps 8 --warmup 3 --no-cpu --no-configs --no-e2e"
timeout 150 $TR --nproc-per-node 8 --master-port 29801 bench.py --gpus 8 $S > gpurun_out/r2u_8gpu_copy_1.json 2> gpurun_out/r2u_8gpu_copy_1.err; echo rc=$?
EALS_ROUTE_COPY=0 timeout 150 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 $S > gpurun_out/r2u_8gpu_kernel_1.json 2> gpurun_out/r2u_8gpu_kernel_1.err; echo rc=$?
timeout 150 $TR --nproc-per-node 8 --master-port 29803 bench.py --gpus 8 $S > gpurun_out/r2u_8gpu_copy_2.json 2> gpurun_out/r2u_8gpu_copy_2.err; echo rc=$?
BEST=$(python - <<'PY'
import json
def ms(f):
    try: return json.loads(open(f).read().strip().split('\n')[-1])['ms_per_step']
    except Exception: return 1e9
c=min(ms('gpurun_out/r2u_8gpu_copy_1.json'), ms('gpurun_out/r2u_8gpu_copy_2.json')); k=ms('gpurun_out/r2u_8gpu_kernel_1.json')
print('1' if c <= k else '0')
PY
)
echo "best EALS_ROUTE_COPY=$BEST"
EALS_ROUTE_COPY=$BEST timeout 240 $TR --nproc-per-node 8 --master-port 29804 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2u_8gpu_full_best$BEST.json 2> gpurun_out/r2u_8gpu_full.err; echo full rc=$?
EALS_ROUTE_COPY=$BEST timeout 150 $TR --nproc-per-node 4 --master-port 29805 bench.py --gpus 4 $S > gpurun_out/r2u_4gpu_best$BEST.json 2> gpurun_out/r2u_4gpu.err; echo n4 rc=$?
EALS_ROUTE_COPY=$BEST timeout 200 python tools/group_bench.py c4 8 8 > gpurun_out/r2u_group_8gpu_best$BEST.json 2> gpurun_out/r2u_group_8gpu.err; echo group rc=$?
