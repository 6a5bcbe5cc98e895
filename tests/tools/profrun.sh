# experiment helper: GPU tests, debug bench (free() time), ncu full capture of the evaluation kernel for both
# workloads, the launch list of the default bench command, compute-sanitizer on a few parity cases
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in json apache; do
  FLBGPU_BENCH_DEBUG=1 timeout 200 python bench.py --workload $wl --primary-only --steps 3 --warmup 3 2> gpurun_out/dbg_$wl.err | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('$wl value %.1f e2e %.1f kms %s total_ms %s' % (d['value']/1e6, d['e2e']['value']/1e6, {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()}, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
  grep "free()" gpurun_out/dbg_$wl.err
  FLBGPU_SLICE_MB=2048 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_chain_eval -s 1 -c 1 -o gpurun_out/r01c_eval_$wl python bench.py --workload $wl --primary-only --lines 1000000 --steps 1 --warmup 1 > gpurun_out/ncu_$wl.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --primary-only > gpurun_out/ncu_launch.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "north_star or json_chain_config1 or l2m_gpu or speculation or sliced" > gpurun_out/r01_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/r01_sanitizer.log
