#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02_gputest_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_8.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_gputest_8.log | tail -12
