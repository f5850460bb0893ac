#!/bin/bash
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2n_bench_n1.json
grep -E "epoch trace|finish trace|gc trace" gpurun_out/r2n_bench_n1.err | tail -4 | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2n_bench_n1b.json 2> gpurun_out/r2n_bench_n1b.err
python -c "$LAST; print('N=1 (no trace): ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2n_bench_n1b.json
FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --workload resnet18 --steps 10 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2n_bench_r18.json 2> gpurun_out/r2n_bench_r18.err
python -c "$LAST; print('r18: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2n_bench_r18.json
grep -E "epoch trace|finish trace|gc trace" gpurun_out/r2n_bench_r18.err | tail -4 | cut -c1-400
timeout 300 python bench.py --workload resnet50x4 --steps 10 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2n_bench_r50.json 2> gpurun_out/r2n_bench_r50.err
python -c "$LAST; print('r50x4: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2n_bench_r50.json
