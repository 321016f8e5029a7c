#!/bin/bash
# Round 2, call 47 (1 GPU): the driver's round-end sequence with the final code -- GPU tests, smoke, bench (both arms)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for i in 1 2; do
echo "factor: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60)"
done
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/call47_tests.log 2>&1; tail -5 gpurun_out/call47_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-1500
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json | cut -c1-400
