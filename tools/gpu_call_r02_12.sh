#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SPRAL_B200_SPLIT=1 SPRAL_B200_SPLIT_TIMEOUT=10 SPRAL_B200_TRACE=1 SPRAL_B200_TRACE_PANELS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n4_split_trace.json 2> gpurun_out/bench_n4_split_trace.err
grep "m 16349" gpurun_out/bench_n4_split_trace.err | tail -64 | sed -e 's/.*p0 \([0-9]*\) done.*urgent \([0-9]*\) bulk \([0-9+]*\) swap \([0-9]*\) *\([0-9.]*\) us since.*first launch to the snapshot, of which \([0-9.]*\) us.*/p0=\1 urgent=\2 bulk=\3 t=\5 wait=\6/' | head -64
