#!/bin/bash
# lna_f32_rows with TMA-staged scores: bytes against the other LNA kernels, then the config-2 step with / without
mkdir -p gpurun_out
AKUGPU_LNA_TMA=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -q -x -k "lna or launch_shape or baseline or config" 2>&1 | tail -2
for rep in 1 2; do
for t in 0 1; do
  AKUGPU_LNA_TMA=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/exp_lnatma$t.json 2> gpurun_out/exp_lnatma$t.err
  python - <<PY
import json
d = json.load(open("gpurun_out/exp_lnatma$t.json"))
print("LNA by TMA $t: %.2f M frames/s, %.2f ms/step, stage_ms %s" % (d["value"] / 1e6, d["ms_per_step"], {k: round(v / 10, 2) for k, v in d["roofline"]["stage_ms"].items()}))
PY
done
done
