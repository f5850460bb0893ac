python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 10 > gpurun_out/f2_default.json 2> gpurun_out/f2_default.err; echo "bench rc=$?"
FRL_B200_FUSE_RELU=1 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/f2_fuse.json 2> gpurun_out/f2_fuse.err
FRL_B200_FUSE_RELU=1 FRL_B200_INPUT_WIRE=bf16 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/f2_fuse_wire.json 2> gpurun_out/f2_fuse_wire.err
python bench.py --steps 30 --warmup 10 --algo adam --no-cpu-baseline --profile gpurun_out/trace_adam.json > gpurun_out/f2_adam.json 2> gpurun_out/f2_adam.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/f2_*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); e=d['e2e']; print(f, round(d['value']), round(d['ms_per_step'],4), 'k2', round(d['roofline']['frac'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), e.get('input_wire'))
PY
