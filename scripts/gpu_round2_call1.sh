#!/bin/bash
# first GPU call of round 2: the whole GPU suite, compute-sanitizer over the small / edge / other-shapes tests, one bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r02_gputest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_1.log
K='small or edge or other_shapes'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/r02_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > gpurun_out/r02_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02_sanitizer_${tool}_pytest.log
done
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_call1.json 2> gpurun_out/r02_bench_call1.err
tail -5 gpurun_out/r02_gputest_1.log
for tool in memcheck racecheck synccheck; do tail -3 gpurun_out/r02_sanitizer_$tool.log; tail -2 gpurun_out/r02_sanitizer_${tool}_pytest.log; done
