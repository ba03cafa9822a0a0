#!/bin/bash
mkdir -p gpurun_out
AKUGPU_BENCH_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_traceA.json 2> gpurun_out/r02_bench_traceA.err
AKUGPU_BENCH_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_traceB.json 2> gpurun_out/r02_bench_traceB.err
echo A; grep -E "produce|config 4|config 5" gpurun_out/r02_bench_traceA.err | head -40
echo B; grep -E "produce|config 4|config 5" gpurun_out/r02_bench_traceB.err | head -40
