#!/bin/bash
# 1 GPU: final check — full GPU suite, smoke, the default bench line
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2t_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2t_pytest.log | tail -1
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])" < gpurun_out/r2t_bench_n1.json
