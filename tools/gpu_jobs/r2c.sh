#!/bin/bash
# 1 GPU: full GPU suite (complete log), K6 tuning sweep, mt / K5 micro-bench, e2e input-path sweep
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2c_pytest.log 2>&1
for r in 1 2 4; do for c in 4 8; do
    echo "ROWS=$r CTAS=$c"
    FRL_B200_COLSUM_ROWS=$r FRL_B200_COLSUM_CTAS=$c timeout 100 python tools/kernel_bench.py --only k6 2>&1 | cut -c1-130
done; done > gpurun_out/r2c_k6_sweep.log 2>&1
timeout 200 python tools/kernel_bench.py --only k2mt,k5 --json gpurun_out/r2c_kernel_bench.json > gpurun_out/r2c_kernel_bench.log 2>&1
for cfg in "kernel 16" "kernel 8" "kernel 4" "kernel 2" "tma 1" "tma 2"; do
    set -- $cfg
    echo "== path $1 blocks $2"
    FRL_B200_EPOCH_TRACE=1 FRL_B200_INPUT_PATH=$1 FRL_B200_INPUT_BLOCKS=$2 timeout 200 python bench.py --steps 20 --warmup 5 \
        --no-cpu-baseline --no-torch-baseline 2> gpurun_out/r2c_e2e_$1_$2.err \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
    grep -E "epoch trace|loader trace" gpurun_out/r2c_e2e_$1_$2.err | tail -4
done > gpurun_out/r2c_e2e_sweep.log 2>&1
tail -15 gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_k6_sweep.log gpurun_out/r2c_kernel_bench.log gpurun_out/r2c_e2e_sweep.log | cut -c1-330
