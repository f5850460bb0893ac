#!/bin/bash
# 2 GPUs: multi-rank parity vs the CPU oracle (world 2), bench N=2 (parity_check inside), K7 launch-shape
# sweep at world 2, stock-PyTorch arm N=2
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/r2e_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2e_pytest.log | tail -2
timeout 100 python tools/kernel_bench.py --only k4,k6 > gpurun_out/r2e_kernel_bench.log 2>&1
cut -c1-150 gpurun_out/r2e_kernel_bench.log
FRL_B200_EPOCH_TRACE=1 timeout 200 python bench.py --steps 20 --warmup 5 --profile gpurun_out/r2e_profile_mlp_b200.json > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
python -c "import json; d=json.load(open('gpurun_out/r2e_bench_n1.json')); print('N=1: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'torch', d['torch_gpu_baseline']['ms_per_step'], 'roof', d['roofline']['frac'])"
grep -E "epoch trace|loader trace" gpurun_out/r2e_bench_n1.err | tail -5 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tests/run_ddp_vs_oracle.py > gpurun_out/r2e_ddp_parity_w2.log 2>&1
echo "ddp parity world 2: $(grep -c DDP_PARITY_OK gpurun_out/r2e_ddp_parity_w2.log) rows ok"
timeout 300 $TR --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2e_bench_n2.json"))
    print("N=2 default: ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity_check"] and d["parity_check"]["ok"],
          "torch", d["torch_gpu_baseline"] and d["torch_gpu_baseline"].get("ms_per_step"))
except Exception as e:
    print("N=2 default bench failed:", e)
PY
for cfg in "16 16 0" "32 16 0" "74 16 0" "74 4 0" "32 8 0" "148 8 1" "74 16 1"; do
    set -- $cfg
    FRL_B200_NVLS_BLOCKS=$1 FRL_B200_NVLS_INFLIGHT=$2 FRL_B200_NVLS_SPLIT_SYNC=$3 timeout 200 $TR --master-port 29603 bench.py --gpus 2 --steps 30 --warmup 5 \
        --no-e2e --no-torch-baseline --no-parity-check 2> gpurun_out/r2e_sweep_$1_$2_$3.err \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks $1 inflight $2 split $3: ms/step', round(d['ms_per_step'],4), 'k7 avg launch ms', d['roofline']['avg_launch_ms'])"
done > gpurun_out/r2e_sweep.log 2>&1
cat gpurun_out/r2e_sweep.log
tail -3 gpurun_out/r2e_ddp_parity_w2.log | cut -c1-200
tail -5 gpurun_out/r2e_bench_n2.err | cut -c1-200
