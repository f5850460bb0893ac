python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 10 > gpurun_out/f3_default.json 2> gpurun_out/f3_default.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f3_*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); e=d['e2e']; print(f, round(d['value']), round(d['ms_per_step'],4), 'k2', round(d['roofline']['frac'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), e.get('input_path'), e.get('input_blocks'))
PY
