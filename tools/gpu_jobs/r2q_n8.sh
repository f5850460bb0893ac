#!/bin/bash
# 8 GPUs: default bench line with the tail split + e2e epoch trace, split on/off comparison,
# multi-rank parity vs the CPU oracle with the split forced onto the toy model
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
FRL_B200_EPOCH_TRACE=1 timeout 400 $TR --master-port 29703 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err
python -c "$LAST; print('N=8: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'parity', d['parity_check']['ok'], d['parity_check']['master'], 'k7 ms', d['roofline']['avg_launch_ms'])" < gpurun_out/r2q_bench_n8.json
grep -E "epoch trace|finish trace" gpurun_out/r2q_bench_n8.err | tail -3 | cut -c1-420
for split in 0 1; do
FRL_B200_TAIL_SPLIT=$split timeout 300 $TR --master-port 2971$split bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2q_bench_n8_s$split.err \
  | python -c "$LAST; print('N=8 split=$split K=40: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'max', d['step_ms_max'])"
done
FRL_B200_TAIL_SPLIT_MIN_BYTES=0 timeout 300 $TR --master-port 29702 tests/run_ddp_vs_oracle.py > gpurun_out/r2q_ddp_parity_world8_split.log 2>&1
echo "ddp parity world 8 (tail split forced): $(grep -c DDP_PARITY_OK gpurun_out/r2q_ddp_parity_world8_split.log) rows ok"; grep -i "split\|error\|FAIL" gpurun_out/r2q_ddp_parity_world8_split.log | head -5
