#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -q -x -s > gpurun_out/r02_gputest_stream.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_stream.log
timeout 300 python scripts/diag_frontend_calls.py > gpurun_out/r02_diag_frontend.log 2>&1
timeout 600 python -m pytest tests/test_gpu_host.py tests/test_gpu_baseline_shapes.py -m gpu -q -k "cpp_tool_flags or config5 or config1" > gpurun_out/r02_gputest_3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_3.log
grep -E "passed|failed|FAILED|ERROR|us per call|stream probe|rc=" gpurun_out/r02_gputest_stream.log | tail -20
cat gpurun_out/r02_diag_frontend.log
tail -3 gpurun_out/r02_gputest_3.log
