#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gputest_14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_14.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_gputest_14.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
tail -12 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
