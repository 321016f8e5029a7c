#!/bin/bash
# near launches split over more CTAs; 64 right-hand sides in one pass again
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q 2>&1 | tail -3
for la in 0 1; do for nr in 1 4 16 64; do
echo "lookahead=$la nrhs=$nr: $(SPRAL_B200_SOLVE_LOOKAHEAD=$la SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done; done
for nr in 1 64; do
SPRAL_B200_TRACE_SOLVE=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr > gpurun_out/solve_tl_$nr.out 2> gpurun_out/solve_tl_$nr.log
done
