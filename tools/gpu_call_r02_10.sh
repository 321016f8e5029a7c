#!/bin/bash
# Round 2, call 10 (2 GPUs): distributed top front with the owner's share; cfg4 at N = 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for split in 1; do
echo "== bench N=2, SPLIT=$split"
SPRAL_B200_SPLIT=$split SPRAL_B200_SPLIT_TIMEOUT=10 SPRAL_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n2_split${split}b.json 2> gpurun_out/bench_n2_split${split}b.err
tail -1 gpurun_out/bench_n2_split${split}b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'])"
grep "trace r\|\[split\]" gpurun_out/bench_n2_split${split}b.err | tail -12
done
echo "== cfg4 at N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --workload cfg4 > gpurun_out/bench_cfg4_n2.json 2> gpurun_out/bench_cfg4_n2.err
tail -1 gpurun_out/bench_cfg4_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'], d['solve_ms'])"
tail -2 gpurun_out/bench_cfg4_n2.err
