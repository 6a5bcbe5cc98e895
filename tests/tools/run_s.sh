WLS="json apache" bash tests/tools/evalvariants.sh FLBGPU_DUMMY=1 FLBGPU_EVAL_PAD_KB=60 FLBGPU_EVAL_PAD_KB=100 > gpurun_out/r02_evalvariants9.txt 2>&1; cat gpurun_out/r02_evalvariants9.txt
