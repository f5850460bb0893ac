python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_bench.csv &
SMI=$!
python bench.py --steps 50 --warmup 10 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 50 --warmup 10 --algo adam --no-cpu-baseline > gpurun_out/final_bench_n1_adam.json 2> gpurun_out/final_bench_n1_adam.err
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 3 --warmup 4 --no-cpu-baseline --graph 0 > gpurun_out/ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:update_kernel -s 3 -c 1 -o gpurun_out/r1b_k2 -f python bench.py --steps 3 --warmup 4 --no-cpu-baseline --no-e2e --graph 0 > gpurun_out/ncu_k2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:criteria_fwd_kernel -s 3 -c 1 -o gpurun_out/r1b_k4 -f python bench.py --steps 3 --warmup 4 --no-cpu-baseline --no-e2e --graph 0 > gpurun_out/ncu_k4.log 2>&1
ls -la gpurun_out/*.ncu-rep
