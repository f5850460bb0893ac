#!/bin/bash
# 4 GPUs: tail split on/off x grid of the last K7 launch, then the default line
N=4
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29820
for cfg in "0 0" "1 0" "0 64" "1 64"; do
set -- $cfg; port=$((port+1))
FRL_B200_TAIL_SPLIT=$1 FRL_B200_NVLS_TAIL_BLOCKS=$2 timeout 300 $TR --master-port $port bench.py --gpus $N --steps 40 --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2s_bench_n${N}_s$1_t$2.err \
  | python -c "$LAST; print('N=$N split=$1 tail=$2 K=40: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'][:3], 'max', d['step_ms_max'])"
done
FRL_B200_EPOCH_TRACE=1 timeout 400 $TR --master-port 29830 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2s_bench_n4.json 2> gpurun_out/r2s_bench_n4.err
python -c "$LAST; print('N=4 default: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'parity', d['parity_check']['ok'], d['parity_check']['master'])" < gpurun_out/r2s_bench_n4.json
grep -E "finish trace: .* flush" gpurun_out/r2s_bench_n4.err | tail -4 | cut -c1-200
