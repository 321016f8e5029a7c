#!/bin/bash
# 2 GPUs: the multi-process GPU tests, cfg5 and cfg4 bench lines with the current kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for wl in cfg5 cfg4; do
echo "== bench $wl N=2"
SPRAL_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_n2.json 2> gpurun_out/bench_${wl}_n2.err
tail -1 gpurun_out/bench_${wl}_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'], d['solve_ms'])"
grep "trace r.* e8" gpurun_out/bench_${wl}_n2.err | sort -t+ -k2 -n | tail -10
done
echo "== bench cfg4 N=1"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 2 --no-cpu-baseline --workload cfg4 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err
tail -1 gpurun_out/bench_cfg4_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'], d['solve_ms'])"
