ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_b.log 2>&1
tail -c 300 gpurun_out/r2h_b.log
