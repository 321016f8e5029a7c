#!/bin/bash
# final bench line of round 2 (N = 1)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
