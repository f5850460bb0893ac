#!/bin/bash
# usage: r2r_tail.sh N "tail grids"  — sweep the grid of the last (exposed) K7 launch, then the default line
N=$1; TAILS=$2
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
port=29800
for tail in $TAILS; do
port=$((port+1))
FRL_B200_NVLS_TAIL_BLOCKS=$tail timeout 300 $TR --master-port $port bench.py --gpus $N --steps 40 --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2r_bench_n${N}_t$tail.err \
  | python -c "$LAST; print('N=$N tail=$tail K=40: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'max', d['step_ms_max'])"
done
