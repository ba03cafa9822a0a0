#!/bin/bash
mkdir -p gpurun_out
for t in 2 3 4; do
  AKUGPU_LNA_TMA=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/exp_lnast$t.json 2> gpurun_out/exp_lnast$t.err
  python - <<PY
import json
d = json.load(open("gpurun_out/exp_lnast$t.json"))
print("LNA ring depth $t: %.2f M frames/s, %.2f ms/step, stage_ms %s" % (d["value"] / 1e6, d["ms_per_step"], {k: round(v / 10, 2) for k, v in d["roofline"]["stage_ms"].items()}))
PY
done
