#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_refbin.py -m gpu -q -s -k "feature or fused or warp_fft or sweep or refbin or goldens or vtln or speaker or pre" > gpurun_out/r02_gputest_11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_11.log
grep -E "passed|failed|FAILED|ERROR|identical:|rc=" gpurun_out/r02_gputest_11.log | tail -20
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_new.json 2> gpurun_out/r02_bench_c3_new.err; grep "config 3" gpurun_out/r02_bench_c3_new.err
AKUGPU_FE_OLDFFT=1 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_old.json 2> gpurun_out/r02_bench_c3_old.err; echo OLD; grep "config 3" gpurun_out/r02_bench_c3_old.err
