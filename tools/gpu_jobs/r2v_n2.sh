#!/bin/bash
# 2 GPUs: grid of the last K7 launch with the tail split on / off, then the default line
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
port=29920
for cfg in "1 0" "1 148" "0 148"; do
set -- $cfg; port=$((port+1))
FRL_B200_TAIL_SPLIT=$1 FRL_B200_NVLS_TAIL_BLOCKS=$2 timeout 300 $TR --master-port $port bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2v_bench_n2_s$1_t$2.err \
  | python -c "$LAST; print('N=2 split=$1 tail=$2 K=40: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'][:3], 'max', d['step_ms_max'])"
done
