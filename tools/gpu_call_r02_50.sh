#!/bin/bash
# inverse diagonal blocks in the T kernels of the large fronts
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -x -q -m gpu 2>&1 | tail -3
for v in 0 1; do for nr in 1 16 64; do
echo "LINV=$v nrhs=$nr: $(SPRAL_B200_SOLVE_LINV=$v SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done; done
