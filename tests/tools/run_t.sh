(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02_gpu_tests_t.log; tail -4 gpurun_out/r02_gpu_tests_t.log
WLS="json" bash tests/tools/evalvariants.sh FLBGPU_DUMMY=1 > gpurun_out/r02_evalvariants10.txt 2>&1; cat gpurun_out/r02_evalvariants10.txt
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02h_launches_json.csv python bench.py --steps 2 --warmup 1 --primary-only --workload json --lines 4000000 > /dev/null) 2>&1 | tail -2
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02h_launches_json.csv') if not l.startswith('=='))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0]
    try: v=float(r[vi].replace(',',''))
    except: continue
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:8]:
    print('  %-50s %4d  %10.1f us total  %9.1f us avg'%(n[:50],c,t/1e3,t/1e3/c))
PY
