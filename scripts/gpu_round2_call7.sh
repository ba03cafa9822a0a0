#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -q -s -k full_size > gpurun_out/r02_gputest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_stream.log
grep -E "passed|failed|FAILED|ERROR|us per|stream probe|rc=" gpurun_out/r02_gputest_stream.log | tail -20
