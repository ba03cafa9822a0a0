#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "launch_shape or chunk or properties" 2>&1 | tail -3
for ch in 0 9472 4736 3712 2432; do
  AKUGPU_CHUNK_FRAMES=$ch python bench.py --steps 6 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02_chunk_$ch.json 2> gpurun_out/r02_chunk_$ch.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_chunk_$ch.json").read())
r=d["roofline"]
print("chunk=$ch: value %.2f M frames/s, ms/step %.2f, e2e %.2f M, scorer launches %d avg %.3f ms, stages %s, clocks %s launches %d" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, r["launches"], r["avg_launch_ms"], {k:round(v/d["steps"],2) for k,v in r["stage_ms"].items()}, d["clocks"]["sm_mhz"], d["gpu_launches"]))
PY
done
