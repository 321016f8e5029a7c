#!/bin/bash
# Round 2, call 35 (8 GPUs): cfg5 and cfg4 at N = 8 and N = 4 with the defaults (proportional mapping, split root front, 3 helpers)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for n in 8 4; do
for wl in cfg5 cfg4; do
echo "== bench $wl N=$n"
SPRAL_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $n --steps 3 --warmup 2 --no-cpu-baseline --workload $wl > gpurun_out/bench_${wl}_n${n}.json 2> gpurun_out/bench_${wl}_n${n}.err
tail -1 gpurun_out/bench_${wl}_n${n}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'], d['solve_ms'])"
grep "trace r.* e8" gpurun_out/bench_${wl}_n${n}.err | sort -t+ -k2 -n | tail -16
done
done
