#!/bin/bash
# 8 GPUs: the default line after the read-back / alignment / tail-grid changes (stock-PyTorch arm: profiles/r2q_bench_n8.json)
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
FRL_B200_EPOCH_TRACE=1 timeout 300 $TR --master-port 29903 bench.py --gpus 8 --steps 20 --warmup 5 --no-torch-baseline > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err
python -c "$LAST; print('N=8: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check']['ok'], d['parity_check']['master'], 'k7 ms', d['roofline']['avg_launch_ms'])" < gpurun_out/r2u_bench_n8.json
grep -E "epoch trace" gpurun_out/r2u_bench_n8.err | tail -2 | cut -c1-420
