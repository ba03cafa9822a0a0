#!/bin/bash
# experiment: what would fewer L2 -> SM bytes of B' buy (gmm_tc16_kernel fetching 1 / 2 of its 3 k-blocks per tile; wrong numbers)
mkdir -p gpurun_out
for rep in 1 2; do
for kb in 0 1 2; do
  AKUGPU_TC16_EXP_KB=$kb timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/exp_kb$kb.json 2> gpurun_out/exp_kb$kb.err
  python - <<PY
import json
d = json.load(open("gpurun_out/exp_kb$kb.json"))
print("k-blocks fetched $kb (0 = all 3): %.2f M frames/s, %.2f ms/step, scorer %.3f ms/launch, clock %s MHz, power %s W" % (
    d["value"] / 1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max")))
PY
done
done
