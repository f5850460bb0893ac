N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29512 tests/run_ddp_vs_oracle.py > gpurun_out/ddp_parity$N.log 2>&1; echo "parity ok count: $(grep -c DDP_PARITY_OK gpurun_out/ddp_parity$N.log)"
python bench.py --gpus 1 --steps 40 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/s4_n1.json 2> gpurun_out/s4_n1.err
$TR --master-port 29514 tools/sweep_nvls.py --bucket-mb 24 --blocks 8,16,37,74 --sync 0 > gpurun_out/sweep${N}_b24.log 2>&1
$TR --master-port 29515 bench.py --gpus $N --steps 30 --warmup 10 > gpurun_out/s4_bench_auto.json 2> gpurun_out/s4_bench_auto.err
FRL_B200_INPUT_PATH=host $TR --master-port 29516 bench.py --gpus $N --steps 30 --warmup 10 > gpurun_out/s4_bench_host.json 2> gpurun_out/s4_bench_host.err
grep -h SWEEP gpurun_out/sweep${N}_b*.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s4_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d.get('e2e') or {}; print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['step_p50_ms'],4), 'e2e', e.get('value'), e.get('ms_per_step'), e.get('input_path'), d['config']['grad_allreduce'][-60:])
PY
