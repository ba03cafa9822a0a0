#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_stream.py -m gpu -q -x -k "cmllr or stream or speaker" > gpurun_out/r02_gputest_16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_16.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " gpurun_out/r02_gputest_16.log | tail -16
