# experiment helper: bench + ncu full capture of the evaluation kernel for both workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in json apache; do
  timeout 200 python bench.py --workload $wl --primary-only --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('$wl value %.1f e2e %.1f kms %s total_ms %s' % (d['value']/1e6, d['e2e']['value']/1e6, {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()}, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
  FLBGPU_SLICE_MB=2048 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_chain_eval -s 1 -c 1 -o gpurun_out/r01b_eval_$wl python bench.py --workload $wl --primary-only --lines 1000000 --steps 1 --warmup 1 > gpurun_out/ncu_$wl.log 2>&1
done
