python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for cfg in kernel:8 kernel:4 tma:2 kernel:16; do
  p=${cfg%%:*}; b=${cfg##*:}
  FRL_B200_INPUT_PATH=$p FRL_B200_INPUT_BLOCKS=$b python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/e2e_${p}_${b}.json 2> gpurun_out/e2e_${p}_${b}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e_kernel_*.json')+glob.glob('gpurun_out/e2e_tma_2.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(f, round(d['value']), d['ms_per_step'], 'e2e', round(e['value']), e['ms_per_step'], e.get('input_path'), e.get('input_blocks'))
PY
