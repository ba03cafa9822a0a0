#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "vtln or sr_norm" > gpurun_out/r02_gputest_17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_17.log
grep -E "passed|failed|FAILED|ERROR|rc=|^E " gpurun_out/r02_gputest_17.log | tail -16
