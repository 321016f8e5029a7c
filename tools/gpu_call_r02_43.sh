#!/bin/bash
# new GPU tests: many right-hand sides (tensor-core solves, 64-wide passes, two lanes), sweeps without look-ahead
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q -k "many_right or paths or agree or switches" 2>&1 | tail -5
