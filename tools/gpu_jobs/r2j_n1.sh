#!/bin/bash
# 1 GPU: full suite, kernel micro-bench after the K4/K6 rewrites, N=1 bench (trace + kernel profiles incl. the e2e epoch),
# ResNet workloads (with kernel-only profiles for both arms)
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2j_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2j_pytest.log | tail -1
timeout 200 python tools/kernel_bench.py --json gpurun_out/r2j_kernel_bench.json > gpurun_out/r2j_kernel_bench.log 2>&1
cut -c1-160 gpurun_out/r2j_kernel_bench.log
FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --profile gpurun_out/r2j_profile_mlp_b200.json > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'max', d['step_ms_max'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'])" < gpurun_out/r2j_bench_n1.json
grep -E "epoch trace|finish trace" gpurun_out/r2j_bench_n1.err | tail -3 | cut -c1-330
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2j_bench_n1_100.json 2> /dev/null
python -c "$LAST; print('100 steps: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2j_bench_n1_100.json
timeout 200 python bench.py --algo adam --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_n1_adam.json 2> /dev/null
python -c "$LAST; print('adam: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'torch', d['torch_gpu_baseline']['ms_per_step'])" < gpurun_out/r2j_bench_n1_adam.json
timeout 200 python bench.py --impl torch-gpu --steps 10 --warmup 5 --no-e2e --profile gpurun_out/r2j_profile_mlp_torch.json > gpurun_out/r2j_torch_n1.json 2> /dev/null
timeout 200 python bench.py --impl torch-gpu --torch-optim fused --steps 20 --warmup 5 --no-e2e > gpurun_out/r2j_torch_n1_fused.json 2> /dev/null
python -c "$LAST; print('torch-gpu fused optimizer: ms/step', d['ms_per_step'])" < gpurun_out/r2j_torch_n1_fused.json
for wl in resnet18 resnet50x4; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --cpu-steps 1 > gpurun_out/r2j_bench_$wl.json 2> gpurun_out/r2j_bench_$wl.err
  python -c "$LAST; print('$wl: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['avg_launch_ms'])" < gpurun_out/r2j_bench_$wl.json
done
timeout 200 python bench.py --workload resnet18 --steps 6 --warmup 4 --graph 0 --no-e2e --no-cpu-baseline --no-torch-baseline --profile gpurun_out/r2j_profile_r18_b200.json > /dev/null 2>&1
timeout 200 python bench.py --impl torch-gpu --workload resnet18 --steps 6 --warmup 4 --no-e2e --profile gpurun_out/r2j_profile_r18_torch.json > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"criteria|colsum" -c 10 -o gpurun_out/r2j_k4k6 python tools/kernel_bench.py --only k4,k6 --iters 1 --warmup 0 --max-sets 2 > gpurun_out/r2j_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/r2j_launches_mlp.csv python bench.py --steps 3 --warmup 4 --graph 0 --no-e2e --no-cpu-baseline --no-torch-baseline > /dev/null 2>&1
