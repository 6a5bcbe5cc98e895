# round 2, re-entry: GPU tests, the default bench line, the launch list of the same command (small), resource usage
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_gpu_tests_x.log; tail -3 gpurun_out/r02_gpu_tests_x.log
(timeout 420 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json) 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
(timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --primary-only --lines 1000000 > gpurun_out/r02_launches_bench.log) 2>&1 | tail -2
