mkdir -p gpurun_out
(timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_ml.csv python bench.py --workload ml --primary-only --steps 1 --warmup 1 --lines 1000000 > /dev/null) 2>&1 | tail -2
