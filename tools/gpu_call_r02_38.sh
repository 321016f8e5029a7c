#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 4 16 64 128; do
echo "nrhs=$nr: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done
for nr in 1 64; do
SPRAL_B200_TRACE_SOLVE=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr > gpurun_out/solve_tl_$nr.out 2> gpurun_out/solve_tl_$nr.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solve or dense or refine" 2>&1 | tail -2
