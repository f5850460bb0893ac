python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 10 --profile gpurun_out/trace5.json > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_loop.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print(round(d['value']), d['ms_per_step'], 'e2e', round(e['value']), e['ms_per_step'], e.get('input_path'), e.get('gpu_launches'), e.get('epoch_losses'))
PY
