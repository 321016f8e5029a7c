#!/bin/bash
# look-ahead inside the wide triangular sweeps; tiles per CTA of the bulk launches
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q 2>&1 | tail -3
for la in 0 1; do for nr in 1 4 16 64; do
echo "lookahead=$la nrhs=$nr: $(SPRAL_B200_SOLVE_LOOKAHEAD=$la SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done; done
run() { echo "$*: $(env "$@" SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 2>&1 | grep '^factor' | cut -c1-60 | tr '\n' ' ')"; }
for t in 1 2 4 1 2 4; do run SPRAL_B200_BULK_TPC=$t; done
