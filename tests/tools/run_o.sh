WLS="json" bash tests/tools/evalvariants.sh "FLBGPU_JSON_BM=1" "FLBGPU_JSON_BM=1 FLBGPU_EVAL_CARVEOUT=30" "FLBGPU_EVAL_CARVEOUT=0" "FLBGPU_EVAL_CARVEOUT=50" > gpurun_out/r02_evalvariants6.txt 2>&1; cat gpurun_out/r02_evalvariants6.txt
cap() { # name workload skip
  (timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chain_eval_t -s $3 -c 1 -o /tmp/$1 python bench.py --steps 1 --warmup 1 --primary-only --workload $2 --lines 1000000 > /dev/null) 2> gpurun_out/r02f_$1.err
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | python tests/tools/ncu_lines.py 150 > gpurun_out/r02f_$1_lines.txt
  head -12 gpurun_out/r02f_$1_lines.txt
}
cap head_json json 2
cap tail_json json 3
cap eval_apache apache 2
ls -la gpurun_out/; du -sh gpurun_out
