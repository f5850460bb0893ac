#!/bin/bash
# 1 GPU: re-run the failing groups with full logs; graph-capture debug; e2e trace
mkdir -p gpurun_out
FRL_B200_DEBUG=1 timeout 400 python -m pytest tests/test_gpu_mlp_parity.py tests/test_gpu_resnet.py tests/test_gpu_kernels.py -m gpu -q --tb=long -s > gpurun_out/r2d_pytest.log 2>&1
timeout 300 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_gpu_mlp_parity.py > gpurun_out/r2d_pytest_rest.log 2>&1
for cfg in "kernel 16" "kernel 8" "tma 2"; do
    set -- $cfg
    echo "== path $1 blocks $2"
    FRL_B200_EPOCH_TRACE=1 FRL_B200_INPUT_PATH=$1 FRL_B200_INPUT_BLOCKS=$2 timeout 200 python bench.py --steps 20 --warmup 5 \
        --no-cpu-baseline --no-torch-baseline 2> gpurun_out/r2d_e2e_$1_$2.err \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('resident', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
    grep -E "epoch trace|loader trace" gpurun_out/r2d_e2e_$1_$2.err | tail -5
done > gpurun_out/r2d_e2e_sweep.log 2>&1
tail -30 gpurun_out/r2d_pytest.log | cut -c1-250
tail -5 gpurun_out/r2d_pytest_rest.log
cat gpurun_out/r2d_e2e_sweep.log | cut -c1-330
