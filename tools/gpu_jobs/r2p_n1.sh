#!/bin/bash
# 1 GPU: full suite + the default bench line + e2e trace after the K8w / read-back changes
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2p_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2p_pytest.log | tail -1
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])" < gpurun_out/r2p_bench_n1.json
grep -E "epoch trace|finish trace" gpurun_out/r2p_bench_n1.err | tail -2 | cut -c1-400
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2p_bench_n1b.json 2> /dev/null
python -c "$LAST; print('N=1 again: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2p_bench_n1b.json
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2p_bench_n1_50.json 2> /dev/null
python -c "$LAST; print('50 steps: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2p_bench_n1_50.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --profile gpurun_out/r2p_profile_mlp_b200.json > /dev/null 2> gpurun_out/r2p_profile.err || tail -3 gpurun_out/r2p_profile.err
