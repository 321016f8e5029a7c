#!/bin/bash
# L2 prefetch of the T kernels' blocks
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for nr in 1 1 4 16 64; do
echo "nrhs=$nr: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done
SPRAL_B200_TRACE_SOLVE=1 SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve 1 > gpurun_out/solve_tl_1.out 2> gpurun_out/solve_tl_1.log
