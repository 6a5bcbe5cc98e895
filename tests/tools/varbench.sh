# experiment helper: bench build variants (fluent-bit_b200/libflbgpu_<V>.so, or the default build) in one GPU call
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in ${VARIANTS:-default}; do
  lib=$PWD/fluent-bit_b200/libflbgpu_$v.so
  [ "$v" = default ] && lib=$PWD/fluent-bit_b200/libflbgpu.so
  for wl in json apache; do
    FLBGPU_LIB=$lib timeout 200 python bench.py --workload $wl --primary-only --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('$v $wl value %.1f e2e %.1f kms %s total_ms %s' % (d['value']/1e6, d['e2e']['value']/1e6, {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()}, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
  done
done
