N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_resnet.py tests/test_gpu_solver.py -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu2.log
$TR --master-port 29515 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2_bench*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d.get('e2e') or {}; print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['step_p50_ms'],4), 'e2e', e.get('value'), e.get('ms_per_step'), e.get('input_path'), json.dumps(d['roofline'])[:600])
PY
tail -3 gpurun_out/s2_bench.err
