#!/bin/bash
# solves after the helper-kernel rewrite; one lane against two lanes of right-hand sides
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solve or dense_fronts or doctored" 2>&1 | tail -3
for lanes in 1 2; do for nr in 1 16 32 64 128; do
echo "lanes=$lanes nrhs=$nr: $(SPRAL_B200_SOLVE_LANES=$lanes SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -2 | tr '\n' ' ')"
done; done
