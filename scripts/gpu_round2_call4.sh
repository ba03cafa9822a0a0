#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -q -s > gpurun_out/r02_gputest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_stream.log
timeout 300 python scripts/diag_config4_calls.py > gpurun_out/r02_diag_config4.log 2>&1
grep -E "passed|failed|FAILED|ERROR|us per call|stream probe|rc=" gpurun_out/r02_gputest_stream.log | tail -20
cat gpurun_out/r02_diag_config4.log
