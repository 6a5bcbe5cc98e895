# round 2, last GPU minutes: the k_l2m_fixup fix and the multiline kernels first, then the rest of the GPU suite, then the multiline bench
mkdir -p gpurun_out
(timeout 150 python -m pytest tests/test_l2m.py tests/test_multiline.py tests/test_shim.py -m gpu -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -12) > gpurun_out/r02_gpu_tests_y1.log; tail -4 gpurun_out/r02_gpu_tests_y1.log
(timeout 150 python -m pytest tests -m gpu -x -q --timeout 60 --timeout-method=thread --ignore tests/test_l2m.py --ignore tests/test_multiline.py --ignore tests/test_shim.py 2>&1 | tail -12) > gpurun_out/r02_gpu_tests_y2.log; tail -4 gpurun_out/r02_gpu_tests_y2.log
(timeout 70 python bench.py --workload ml --primary-only --steps 3 --warmup 3 --lines 4000000 > gpurun_out/r02_bench_ml.json) 2> gpurun_out/r02_bench_ml.err; tail -c 1500 gpurun_out/r02_bench_ml.json; tail -3 gpurun_out/r02_bench_ml.err
