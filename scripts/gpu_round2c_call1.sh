#!/bin/bash
# round 2, third session: full GPU suite, streaming sub-record, sanitizer pass over the kernels added / changed in this session
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_gputest.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/r02c_gputest.log
timeout -s KILL 300 python scripts/stream_sub.py > gpurun_out/r02c_stream_sub.json 2> gpurun_out/r02c_stream_sub.err; echo "stream_sub rc=$?"; tail -2 gpurun_out/r02c_stream_sub.err
SEL="resident_scorer_returns or resident_scorer_lifecycle or fused or warp_fft"
for tool in memcheck synccheck racecheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/r02c_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02c_sanitizer_$tool.log | tail -3
done
