#!/bin/bash
# evaluation-kernel variants on one GPU: device-resident value and kernel ms of the JSON and apache chains
run() { echo "== $*"; for wl in ${WLS:-json apache}; do env "$@" timeout 120 python bench.py --steps 3 --warmup 2 --primary-only --workload $wl --lines 4000000 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   %-6s value %.1f M lines/s  eval %.2f ms  emit %.2f ms  index %.2f ms' % ('$wl', d['value']/1e6, d['kernel_ms_per_step']['evaluate'], d['kernel_ms_per_step']['emit'], d['kernel_ms_per_step']['index']))
"; done; }
for v in "$@"; do run $v; done
