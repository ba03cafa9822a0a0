#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/exp_lna.json 2> gpurun_out/exp_lna.err
python - <<PY
import json
d = json.load(open("gpurun_out/exp_lna.json"))
print("%.2f M frames/s, %.2f ms/step, stage_ms %s" % (d["value"] / 1e6, d["ms_per_step"], d["roofline"]["stage_ms"]))
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "lna" 2>&1 | tail -2
