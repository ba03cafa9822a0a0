#!/bin/bash
# final state of the resident scorer: streaming sub-record + sanitizer over its tests
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/stream_sub.py > gpurun_out/r02c_stream_sub.json 2> gpurun_out/r02c_stream_sub.err; echo "stream_sub rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_stream.py -m gpu -q -x -k "resident_scorer_returns or resident_scorer_lifecycle or session_mirror" \
    > gpurun_out/r02c_sanitizer5_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02c_sanitizer5_$tool.log | tail -3
done
