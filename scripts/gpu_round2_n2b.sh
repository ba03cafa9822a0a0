#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_multigpu.py tests/test_gpu_parity.py -m gpu -q -x -k "launch_shape or checksums or chunk or properties or p2p" > gpurun_out/r02_gputest_n2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_n2b.log
tail -5 gpurun_out/r02_gputest_n2b.log
bash scripts/gpu_round2_n2.sh
