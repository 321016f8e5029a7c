#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "factor: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60)"
( time timeout 1500 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "parity or dense or doctored" 2>&1 | tail -2
