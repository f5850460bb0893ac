#!/bin/bash
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 200 python -m pytest tests/test_gpu_resnet.py -m gpu -q --tb=short > gpurun_out/r2i_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2i_pytest.log | tail -1
FRL_B200_EPOCH_TRACE=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile gpurun_out/r2i_profile_mlp_b200.json > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'max', d['step_ms_max'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2i_bench_n1.json
grep -E "epoch trace|finish trace" gpurun_out/r2i_bench_n1.err | tail -3 | cut -c1-330
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for k in 20 40; do
timeout 300 $TR --master-port 2960$((k/10)) bench.py --gpus 2 --steps $k --warmup 5 --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2i_bench_n2_k$k.err \
  | python -c "$LAST; print('N=2 K=$k: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'max', d['step_ms_max'])"
done
timeout 300 $TR --master-port 29607 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
python -c "$LAST; print('N=2 default: ms/step', d['ms_per_step'], 'p50', d['step_p50_ms'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity_check']['ok'])" < gpurun_out/r2i_bench_n2.json
