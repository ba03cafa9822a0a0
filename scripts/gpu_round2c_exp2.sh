#!/bin/bash
# gmm_tc16_kernel with the fetch of B' shared by clusters of 1 / 2 / 4 frame tiles (tma multicast): config-2 step, alternating
mkdir -p gpurun_out
for rep in 1 2; do
for cl in 1 2 4; do
  AKUGPU_TC16_CLUSTER=$cl timeout 300 python bench.py --steps 10 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/exp_cl$cl.json 2> gpurun_out/exp_cl$cl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/exp_cl$cl.json"))
print("cluster $cl: %.2f M frames/s, %.2f ms/step, scorer %.3f ms/launch, clock %s MHz, power %s W" % (
    d["value"] / 1e6, d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max")))
PY
done
done
