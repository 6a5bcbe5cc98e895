# launch list + per-kernel ncu summaries of the split evaluation (JSON chain), sizes kept small for the copy back
export FLBGPU_EVAL_SPLIT=1
for wl in json apache; do
  (timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches_$wl.csv python bench.py --steps 2 --warmup 1 --primary-only --workload $wl --lines 4000000 > /dev/null) 2>&1 | tail -2
  (timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chain_eval -s 4 -c 2 -o /tmp/ev_$wl python bench.py --steps 1 --warmup 1 --primary-only --workload $wl --lines 1000000 > /dev/null) 2> gpurun_out/r02e_ncu_$wl.err
  ncu -i /tmp/ev_$wl.ncu-rep --page raw --csv > gpurun_out/r02e_eval_${wl}_raw.csv 2>/dev/null
  ncu -i /tmp/ev_$wl.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip -9 > gpurun_out/r02e_eval_${wl}_cuda.csv.gz
done
ls -la gpurun_out/; du -sh gpurun_out
