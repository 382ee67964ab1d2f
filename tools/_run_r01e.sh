timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_r01e_list.log 2>&1
ls -la gpurun_out/
