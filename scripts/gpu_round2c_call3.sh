#!/bin/bash
# the driver's N = 1 command line on the final code of the third session + the reference arm
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02c_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_arm.json 2> gpurun_out/r02c_bench_reference_arm.err; echo "ref rc=$?"
