# experiment helper: A/B of an environment knob on the same box, alternating runs
# usage: KNOB=FLBGPU_TAPER A=1 B=0 bash tests/tools/abbench.sh
mkdir -p gpurun_out
for rep in 1 2 3; do
  for v in "$A" "$B"; do
    for wl in ${WLS:-json apache}; do
      env $KNOB=$v timeout 200 python bench.py --workload $wl --primary-only --steps 3 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('$KNOB=$v $wl value %.1f e2e %.1f total_ms %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e'].get('host_phase_ms_last_call',{}).get('total')))
"
    done
  done
done
