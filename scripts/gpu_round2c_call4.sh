#!/bin/bash
# sanitizer pass over the front-end kernels added after the third pass (fe_spectrum_bfft, fe_delta2_merge_tiled) + resident scorer timing
mkdir -p gpurun_out
SEL="fused or warp_fft or frontend or feature or golden or aku or sweep or resident_scorer_returns"
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/r02c_sanitizer4_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02c_sanitizer4_$tool.log | tail -3
done
timeout 300 python -m pytest tests/test_gpu_stream.py -q -s -k "resident and full_size" 2>&1 | grep "timed inside\|passed\|failed"
