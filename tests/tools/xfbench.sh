# experiment helper: download ring geometry, apache chain (large output) end to end
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "8 16" "16 8" "16 16" "24 8" "32 4"; do
  set -- $cfg
  FLBGPU_XF_SLOTS=$1 FLBGPU_XF_MB=$2 timeout 200 python bench.py --workload apache --primary-only --steps 4 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('slots $1 x $2 MB apache e2e %.1f total_ms %s' % (d['e2e']['value']/1e6, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
done
done
