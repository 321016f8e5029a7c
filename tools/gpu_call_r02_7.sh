#!/bin/bash
# Round 2, call 7 (2 GPUs): the two-rank GPU tests, then bench at N = 2 with the proportional partition and the reference one
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/call7_tests.log 2>&1; tail -5 gpurun_out/call7_tests.log
for mode in proportional reference; do
echo "== bench N=2, partition $mode"
SPRAL_B200_PARTITION=$mode SPRAL_B200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err
tail -1 gpurun_out/bench_n2_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['solve_ms'], d['backward_error'])"
grep "trace r" gpurun_out/bench_n2_$mode.err | tail -8
done
