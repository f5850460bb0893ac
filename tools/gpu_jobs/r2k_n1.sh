#!/bin/bash
# 1 GPU: final verification — full suite, smoke, default bench, ResNet workloads, complete per-kernel lists
mkdir -p gpurun_out
LAST='import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])'
timeout 500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2k_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2k_pytest.log | tail -1
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
python -c "$LAST; print('N=1: ms/step', d['ms_per_step'], 'first5', d['step_ms_first5'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])" < gpurun_out/r2k_bench_n1.json
grep -E "epoch trace|finish trace" gpurun_out/r2k_bench_n1.err | tail -2 | cut -c1-330
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2k_bench_n1_50.json 2> /dev/null
python -c "$LAST; print('50 steps: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" < gpurun_out/r2k_bench_n1_50.json
timeout 200 python bench.py --impl torch-gpu --torch-optim fused --steps 20 --warmup 5 --no-e2e > gpurun_out/r2k_torch_n1_fused.json 2> /dev/null
python -c "$LAST; print('torch-gpu fused optimizer: ms/step', d['ms_per_step'])" < gpurun_out/r2k_torch_n1_fused.json
for wl in resnet18 resnet50x4; do
  FRL_B200_EPOCH_TRACE=1 timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --cpu-steps 1 > gpurun_out/r2k_bench_$wl.json 2> gpurun_out/r2k_bench_$wl.err
  python -c "$LAST; print('$wl: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['avg_launch_ms'])" < gpurun_out/r2k_bench_$wl.json
  grep -E "epoch trace" gpurun_out/r2k_bench_$wl.err | tail -1 | cut -c1-330
done
timeout 200 python bench.py --workload resnet18 --steps 6 --warmup 4 --graph 0 --no-e2e --no-cpu-baseline --no-torch-baseline --profile gpurun_out/r2k_profile_r18_b200.json > /dev/null 2>&1
timeout 200 python bench.py --impl torch-gpu --workload resnet18 --steps 6 --warmup 4 --no-e2e --profile gpurun_out/r2k_profile_r18_torch.json > /dev/null 2>&1
timeout 200 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-torch-baseline --profile gpurun_out/r2k_profile_mlp_b200.json > /dev/null 2>&1
timeout 200 python bench.py --impl torch-gpu --steps 10 --warmup 5 --no-e2e --profile gpurun_out/r2k_profile_mlp_torch.json > /dev/null 2>&1
timeout 100 python tools/kernel_bench.py --only k4,k6 2>&1 | cut -c1-150
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | python -c "$LAST; print('reference arm (unmodified reference, batch 4096):', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
