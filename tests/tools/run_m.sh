(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_gpu_tests_m.log; tail -3 gpurun_out/r02_gpu_tests_m.log
WLS="json apache nginx" bash tests/tools/evalvariants.sh FLBGPU_EVAL_SPLIT=1 FLBGPU_EVAL_SPLIT=0 > gpurun_out/r02_evalvariants5.txt 2>&1; cat gpurun_out/r02_evalvariants5.txt
for wl in json apache; do
  (timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chain_eval -s 4 -c 2 -o /tmp/ev_$wl python bench.py --steps 1 --warmup 1 --primary-only --workload $wl --lines 1000000 > /dev/null) 2> gpurun_out/r02d_ncu_$wl.err
  ncu -i /tmp/ev_$wl.ncu-rep --page raw --csv > gpurun_out/r02d_eval_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/ev_$wl.ncu-rep --page source --csv --print-source sass > gpurun_out/r02d_eval_${wl}_sass.csv 2>/dev/null
done
ls -la gpurun_out/ | tail; du -sh gpurun_out
