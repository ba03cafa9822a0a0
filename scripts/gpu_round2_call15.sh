#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_refbin.py -m gpu -q -s > gpurun_out/r02_gputest_15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_15.log
grep -E "passed|failed|FAILED|ERROR|us per|rc=" gpurun_out/r02_gputest_15.log | tail -24
