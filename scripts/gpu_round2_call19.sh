#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "tolerance or other_shapes or properties or config2 or config4 or launch_shape or hybrid or full_cov or config5" > gpurun_out/r02_gputest_19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_19.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_gputest_19.log | tail -6
for v in new old new old; do
  if [ $v = old ]; then export AKUGPU_TC16_EPI_OLD=1; else unset AKUGPU_TC16_EPI_OLD; fi
  python bench.py --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02_epi_$v.json 2> gpurun_out/r02_epi_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_epi_$v.json").read()); r=d["roofline"]
print("epilogue=$v: value %.2f M frames/s, ms/step %.2f, scorer avg launch %.3f ms, issued frac %.3f, mufu frac %.3f, clocks %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["issued_mma"]["frac_of_peak"], r["mufu_view"]["frac"], d["clocks"]["sm_mhz"]))
PY
done
