(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02_gpu_tests_r.log; tail -4 gpurun_out/r02_gpu_tests_r.log
(timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_r_n1.json) 2> gpurun_out/r02_bench_r.err; tail -c 300 gpurun_out/r02_bench_r.err; head -c 600 gpurun_out/r02_bench_r_n1.json; echo
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_r.json) 2> gpurun_out/r02_ref_r.err; head -c 300 gpurun_out/r02_ref_r.json; echo
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --primary-only > gpurun_out/r02_launches_run.log) 2>&1 | tail -2
du -sh gpurun_out
