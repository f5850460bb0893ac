python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 10 --profile gpurun_out/trace3.json > gpurun_out/bench_host.json 2> gpurun_out/bench_host.err; echo "bench rc=$?"
for t in 8 32; do FRL_B200_INPUT_THREADS=$t python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_host_t$t.json 2> gpurun_out/bench_host_t$t.err; done
FRL_B200_INPUT_PATH=tma FRL_B200_INPUT_BLOCKS=2 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_tma2.json 2> gpurun_out/bench_tma2.err
FRL_B200_INPUT_PATH=kernel FRL_B200_INPUT_BLOCKS=8 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_k8.json 2> gpurun_out/bench_k8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_host*.json')+glob.glob('gpurun_out/bench_tma2.json')+glob.glob('gpurun_out/bench_k8.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(f, round(d['value']), d['ms_per_step'], 'e2e', round(e['value']), e['ms_per_step'], e.get('input_path'), e.get('input_threads'))
PY
