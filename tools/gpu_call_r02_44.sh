#!/bin/bash
# half-height T kernels on the levels of small fronts: memcheck of the many-right-hand-sides tests, timings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 compute-sanitizer --error-exitcode 9 --print-limit 10 --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "many_right and (16 or 64)" > gpurun_out/sanitizer44.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer44.log | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -q -m gpu -x -k "many_right or solve or paths or agree or switches" 2>&1 | tail -2
for nr in 1 16 32 64 128; do
echo "nrhs=$nr: $(SPRAL_B200_NOPROFILE=1 timeout 600 python tools/profile_factor.py 100 indef solve $nr 2>&1 | grep 'solve nrhs' | tail -1)"
done
