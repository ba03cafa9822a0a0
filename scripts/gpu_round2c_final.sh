#!/bin/bash
# final check of the third session: smoke, the whole GPU suite, the driver's N = 1 command line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_gputest_final.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/r02c_gputest_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1_final.json 2> gpurun_out/r02c_bench_n1_final.err; echo "bench rc=$?"
head -c 400 gpurun_out/r02c_bench_n1_final.json
