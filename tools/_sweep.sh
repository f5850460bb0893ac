python -m pytest tests/test_gpu_resnet.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu3.log
python bench.py --steps 50 --warmup 10 --algo adam --no-cpu-baseline > gpurun_out/final_bench_n1_adam.json 2> gpurun_out/final_bench_n1_adam.err
python - <<'PY'
import json
for l in open('gpurun_out/final_bench_n1_adam.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print(round(d['value']), d['ms_per_step'], 'e2e', round(e['value']), e['ms_per_step'], d['roofline']['frac'])
PY
