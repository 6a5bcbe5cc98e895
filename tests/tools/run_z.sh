# round 2, the build with the second-byte filter of the regex search: whole GPU suite once, then the multiline bench again
mkdir -p gpurun_out
(timeout 105 python -m pytest tests -m gpu -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -8) > gpurun_out/r02_gpu_tests_z.log; tail -3 gpurun_out/r02_gpu_tests_z.log
(timeout 45 python bench.py --workload ml --primary-only --steps 3 --warmup 3 --lines 4000000 > gpurun_out/r02_bench_ml2.json) 2> gpurun_out/r02_bench_ml2.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/r02_bench_ml2.json"))
    print("ml value %.1f M lines/s, e2e %.1f M, kernel ms %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["kernel_ms_per_step"]))
except Exception as e:
    print("no bench line:", e)
P
tail -2 gpurun_out/r02_bench_ml2.err
