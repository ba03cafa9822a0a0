#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -q -s -k "stream or fused or feature" > gpurun_out/r02_gputest_13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_13.log
grep -E "passed|failed|FAILED|ERROR|us per|stream probe|rc=" gpurun_out/r02_gputest_13.log | tail -24
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_new.json 2> gpurun_out/r02_bench_c3_new.err; grep "config 3" gpurun_out/r02_bench_c3_new.err | head -3
python bench.py --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02_bench_head.json 2> gpurun_out/r02_bench_head.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_head.json").read()); r=d["roofline"]
print("headline: value %.2f M frames/s, ms/step %.2f, e2e %.2f M, stages/step %s, launches %d" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, {k:round(v/d["steps"],2) for k,v in r["stage_ms"].items()}, d["gpu_launches"]))
PY
