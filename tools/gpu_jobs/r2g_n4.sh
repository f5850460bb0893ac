#!/bin/bash
# 4 GPUs: multi-rank parity vs the CPU oracle at world 4, bench N=4 (parity_check inside), ResNet-18 N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29701 tests/run_ddp_vs_oracle.py > gpurun_out/r2g_ddp_parity_w4.log 2>&1
echo "ddp parity world 4: $(grep -c DDP_PARITY_OK gpurun_out/r2g_ddp_parity_w4.log) rows ok"
timeout 300 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 20 --warmup 5 \
    --profile gpurun_out/r2g_profile_mlp_b200_n4.json > gpurun_out/r2g_bench_n4.json 2> gpurun_out/r2g_bench_n4.err
timeout 300 $TR --nproc-per-node 4 --master-port 29706 bench.py --gpus 4 --workload resnet18 --steps 10 --warmup 5 \
    > gpurun_out/r2g_bench_r18_n4.json 2> gpurun_out/r2g_bench_r18_n4.err
timeout 300 $TR --nproc-per-node 4 --master-port 29705 bench.py --impl torch-gpu --gpus 4 --steps 20 --warmup 5 \
    --profile gpurun_out/r2g_profile_mlp_torch_n4.json > gpurun_out/r2g_torch_n4.json 2> gpurun_out/r2g_torch_n4.err
python - <<'PY'
import json
def load(p):
    return json.loads([l for l in open(p).read().splitlines() if l.startswith("{")][-1])
for name in ("bench_n4", "bench_r18_n4", "torch_n4"):
    try:
        d = load("gpurun_out/r2g_%s.json" % name)
        e = d.get("e2e") or {}
        print(name, "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), "first5", d.get("step_ms_first5"), "e2e ms", e.get("ms_per_step"),
              "parity", (d.get("parity_check") or {}).get("ok"), "torch ms", (d.get("torch_gpu_baseline") or {}).get("ms_per_step"))
    except Exception as ex:
        print(name, "FAILED", ex)
PY
tail -2 gpurun_out/r2g_ddp_parity_w4.log | cut -c1-200
tail -4 gpurun_out/r2g_bench_n4.err | cut -c1-250
