#!/bin/bash
# Round 2, call 8 (2 GPUs): distributed top front on hardware for the first time
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for split in 0 1; do
echo "== bench N=2, SPLIT=$split"
SPRAL_B200_SPLIT=$split SPRAL_B200_SPLIT_TIMEOUT=10 SPRAL_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n2_split$split.json 2> gpurun_out/bench_n2_split$split.err
tail -1 gpurun_out/bench_n2_split$split.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'])"
grep "trace r\|\[split\]" gpurun_out/bench_n2_split$split.err | tail -14
tail -3 gpurun_out/bench_n2_split$split.err
done
