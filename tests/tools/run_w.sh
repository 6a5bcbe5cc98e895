(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_gpu_tests_w.log; tail -3 gpurun_out/r02_gpu_tests_w.log
WLS="json apache nginx" bash tests/tools/evalvariants.sh FLBGPU_DUMMY=1 > gpurun_out/r02_evalvariants13.txt 2>&1; cat gpurun_out/r02_evalvariants13.txt
cap() { # name workload skip
  (timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chain_eval_t -s $3 -c 1 -o /tmp/$1 python bench.py --steps 1 --warmup 1 --primary-only --workload $2 --lines 1000000 > /dev/null) 2> gpurun_out/r02i_$1.err
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02i_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | python tests/tools/ncu_lines.py 120 > gpurun_out/r02i_$1_lines.txt
  head -3 gpurun_out/r02i_$1_lines.txt
}
cap head_json json 2
cap tail_json json 3
