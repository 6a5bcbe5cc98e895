# helper: the driver's multi-GPU launch of bench.py, summarised
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('N=%d value %.1f e2e %.1f ms %.1f north_star %.1f / %.1f' % (d['n_gpus'], d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step'], d['north_star']['value']/1e6, d['north_star']['e2e']/1e6))
"
tail -2 gpurun_out/bench_n$N.err
