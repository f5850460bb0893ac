run() { # name, env..., -- args
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 --no-e2e $EXTRA > gpurun_out/s3_$name.json 2> gpurun_out/s3_$name.err
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/run_ddp_vs_oracle.py > gpurun_out/ddp_parity2.log 2>&1; grep -c DDP_PARITY_OK gpurun_out/ddp_parity2.log
python bench.py --gpus 1 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/s3_n1.json 2> gpurun_out/s3_n1.err
EXTRA="" run old74 FRL_B200_NVLS_BLOCKS=74 FRL_B200_NVLS_SPLIT_SYNC=0
EXTRA="" run split74 FRL_B200_NVLS_BLOCKS=74
EXTRA="" run split148 FRL_B200_NVLS_BLOCKS=148
EXTRA="" run split32 FRL_B200_NVLS_BLOCKS=32
EXTRA="--bucket-mb 24" run split74_b24 FRL_B200_NVLS_BLOCKS=74
EXTRA="--bucket-mb 24" run split148_b24 FRL_B200_NVLS_BLOCKS=148
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s3_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['step_p50_ms'],4), d['roofline']['avg_launch_ms'])
PY
