#!/bin/bash
# 1 GPU: parity re-run, ncu pages for K4/K6/K8/K2-mt, e2e with the re-written gather, ResNet launch lists
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp_parity.py -m gpu -q --tb=short -s > gpurun_out/r2f_pytest_mlp.log 2>&1
timeout 400 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_mlp_parity.py > gpurun_out/r2f_pytest_rest.log 2>&1
timeout 120 python tools/kernel_bench.py --only k8,k4 --json gpurun_out/r2f_kernel_bench.json > gpurun_out/r2f_kernel_bench.log 2>&1
for rows in 1 4; do
  FRL_B200_COLSUM_ROWS=$rows timeout 200 ncu --set full --clock-control none --import-source on -k regex:"colsum" -c 3 \
      -o gpurun_out/r2f_k6_rows$rows python tools/kernel_bench.py --only k6 --iters 1 --warmup 0 --max-sets 2 > gpurun_out/r2f_ncu_k6_$rows.log 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"criteria|gather_rows_kernel|update_mt|flatten_kernel|affine" -c 14 \
    -o gpurun_out/r2f_small python tools/kernel_bench.py --only k4,k5,k8,k2mt --iters 1 --warmup 0 --max-sets 2 > gpurun_out/r2f_ncu_small.log 2>&1
for cfg in "kernel 8" "kernel 4" "tma 2"; do
    set -- $cfg
    echo "== path $1 blocks $2"
    FRL_B200_EPOCH_TRACE=1 FRL_B200_INPUT_PATH=$1 FRL_B200_INPUT_BLOCKS=$2 timeout 200 python bench.py --steps 20 --warmup 5 \
        --no-cpu-baseline --no-torch-baseline 2> gpurun_out/r2f_e2e_$1_$2.err \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
    grep -E "epoch trace|loader trace" gpurun_out/r2f_e2e_$1_$2.err | tail -6
done > gpurun_out/r2f_e2e_sweep.log 2>&1
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/r2f_bench_mlp_200.json 2> gpurun_out/r2f_bench_mlp_200.err
timeout 200 python bench.py --workload resnet18 --steps 6 --warmup 4 --graph 0 --no-e2e --no-cpu-baseline --no-torch-baseline \
    --profile gpurun_out/r2f_profile_r18_b200.json > gpurun_out/r2f_r18_prof.json 2> gpurun_out/r2f_r18_prof.err
timeout 200 python bench.py --impl torch-gpu --workload resnet18 --steps 6 --warmup 4 --no-e2e \
    --profile gpurun_out/r2f_profile_r18_torch.json > gpurun_out/r2f_r18_torch.json 2> gpurun_out/r2f_r18_torch.err
timeout 200 python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline --no-torch-baseline \
    --profile gpurun_out/r2f_profile_mlp_b200.json > /dev/null 2> gpurun_out/r2f_mlp_prof.err
timeout 200 python bench.py --impl torch-gpu --steps 10 --warmup 5 --no-e2e \
    --profile gpurun_out/r2f_profile_mlp_torch.json > gpurun_out/r2f_mlp_torch.json 2> gpurun_out/r2f_mlp_torch.err
tail -40 gpurun_out/r2f_pytest_mlp.log | cut -c1-220
tail -4 gpurun_out/r2f_pytest_rest.log
cat gpurun_out/r2f_kernel_bench.log gpurun_out/r2f_e2e_sweep.log | cut -c1-300
python -c "import json; d=json.load(open('gpurun_out/r2f_bench_mlp_200.json')); print('200 steps: resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
