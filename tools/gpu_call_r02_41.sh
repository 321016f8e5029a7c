#!/bin/bash
# Round 2, call 41 (1 GPU): the driver's round-end sequence with the final kernels -- GPU tests, smoke, bench (both arms) --
# and the ncu launch lists of the bench command and of the solves
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/call41_tests.log 2>&1; tail -6 gpurun_out/call41_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-3000
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json | cut -c1-1200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_bench41.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench41.csv "bench.py --steps 2 --warmup 1 under ncu" | head -30
for nr in 1 64; do
SPRAL_B200_NOPROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
   --log-file gpurun_out/launches_solve41_$nr.csv python tools/profile_factor.py 100 indef solve $nr > gpurun_out/prof_solve41_$nr.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_solve41_$nr.csv "solve nrhs=$nr" 2>/dev/null | head -18
done
gzip -f gpurun_out/launches_bench41.csv gpurun_out/launches_solve41_1.csv gpurun_out/launches_solve41_64.csv
