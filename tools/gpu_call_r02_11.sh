#!/bin/bash
# Round 2, call 11 (8 GPUs): cfg5 at N = 8 and N = 4, proportional partition + distributed root front
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for n in 8 4; do
echo "== bench N=$n, SPLIT=1"
SPRAL_B200_SPLIT=1 SPRAL_B200_SPLIT_TIMEOUT=10 SPRAL_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $n --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n${n}_split1.json 2> gpurun_out/bench_n${n}_split1.err
tail -1 gpurun_out/bench_n${n}_split1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['nparts'], d['inform'], d['backward_error'], d['solve_ms'])"
grep "trace r.* e8\|\[split\]" gpurun_out/bench_n${n}_split1.err | tail -24
done
