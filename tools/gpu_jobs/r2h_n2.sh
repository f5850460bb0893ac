#!/bin/bash
# 2 GPUs: N=1 bench (e2e trace), N=2 default + K7 launch-shape sweep at world 2, failing test re-run
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 300 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_solver.py -m gpu -q --tb=short > gpurun_out/r2h_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2h_pytest.log | tail -1
FRL_B200_EPOCH_TRACE=1 timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'])" < gpurun_out/r2h_bench_n1.json
grep -E "epoch trace" gpurun_out/r2h_bench_n1.err | tail -1 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
python -c "$LAST; print('N=2 default: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check']['ok'], 'torch', d['torch_gpu_baseline']['ms_per_step'])" < gpurun_out/r2h_bench_n2.json
for cfg in "16 16 0" "32 16 0" "74 16 0" "74 4 0" "32 8 0" "148 8 1" "74 16 1" "16 4 0"; do
    set -- $cfg
    FRL_B200_NVLS_BLOCKS=$1 FRL_B200_NVLS_INFLIGHT=$2 FRL_B200_NVLS_SPLIT_SYNC=$3 timeout 200 $TR --master-port 29603 bench.py --gpus 2 --steps 40 --warmup 5 \
        --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2h_sweep_$1_$2_$3.err \
        | python -c "$LAST; print('blocks $1 inflight $2 split $3: ms/step', round(d['ms_per_step'],4), 'p50', round(d['step_p50_ms'],4), 'k7 avg launch ms', d['roofline']['avg_launch_ms'])"
done > gpurun_out/r2h_sweep.log 2>&1
FRL_B200_BUCKET_MB=48 timeout 200 $TR --master-port 29604 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> /dev/null \
    | python -c "$LAST; print('bucket 48 MiB: ms/step', round(d['ms_per_step'],4), 'p50', round(d['step_p50_ms'],4))" >> gpurun_out/r2h_sweep.log 2>&1
cat gpurun_out/r2h_sweep.log
