(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_gpu_tests_v.log; tail -3 gpurun_out/r02_gpu_tests_v.log
WLS="json apache" bash tests/tools/evalvariants.sh FLBGPU_DUMMY=1 FLBGPU_SLICE_MB=256 FLBGPU_SLICE_MB=64 > gpurun_out/r02_evalvariants12.txt 2>&1; cat gpurun_out/r02_evalvariants12.txt
