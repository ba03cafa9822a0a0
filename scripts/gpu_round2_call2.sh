#!/bin/bash
# whole GPU suite (no -x: see everything), then the N=1 bench line with its sub-records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=20 > gpurun_out/r02_gputest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_call2.json 2> gpurun_out/r02_bench_n1_call2.err; echo "bench rc=$?" >> gpurun_out/r02_bench_n1_call2.err
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r02_gputest_2.log | tail -20
tail -15 gpurun_out/r02_bench_n1_call2.err
