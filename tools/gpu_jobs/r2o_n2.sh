#!/bin/bash
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for split in 1 0; do
for k in 20 40; do
FRL_B200_TAIL_SPLIT=$split timeout 300 $TR --master-port 2961$((k/10)) bench.py --gpus 2 --steps $k --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2o_bench_n2_s${split}_k$k.err \
  | python -c "$LAST; print('N=2 split=$split K=$k: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'max', d['step_ms_max'])"
done
done
timeout 300 $TR --master-port 29627 bench.py --gpus 2 --steps 20 --warmup 5 --no-torch-baseline > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err
python -c "$LAST; print('N=2 default: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check'])" < gpurun_out/r2o_bench_n2.json
timeout 300 $TR --master-port 29637 tests/run_ddp_vs_oracle.py > gpurun_out/r2o_ddp_parity_world2.log 2>&1; tail -12 gpurun_out/r2o_ddp_parity_world2.log | cut -c1-200
