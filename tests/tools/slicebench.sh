# experiment helper: end-to-end throughput for several slice sizes (+ plain H2D rate with FLBGPU_BENCH_DEBUG)
for mb in ${SLICES:-32 64 128 256}; do
  for wl in json apache; do
    FLBGPU_BENCH_DEBUG=1 FLBGPU_SLICE_MB=$mb timeout 200 python bench.py --workload $wl --primary-only --steps 3 --warmup 3 2>gpurun_out/dbg.err | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('slice $mb $wl value %.1f e2e %.1f total_ms %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
    grep -E "H2D|free" gpurun_out/dbg.err
  done
done
