#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_multigpu.py -m gpu -q -x -k "overlapped or launch_shape or properties or chunk or device_buffers or checksums" > gpurun_out/r02_gputest_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_9.log
tail -4 gpurun_out/r02_gputest_9.log
for cfg in "1 0" "0 3" "0 2" "1 2"; do
  set -- $cfg
  AKUGPU_OVERLAP=$1 AKUGPU_TC16_SLOTS=$2 python bench.py --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r02_ovl_$1_$2.json 2> gpurun_out/r02_ovl_$1_$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_ovl_$1_$2.json").read())
r=d["roofline"]
print("overlap=$1 slots=$2: value %.2f M frames/s, ms/step %.2f, e2e %.2f M, scorer avg launch %.3f ms, stages %s, clocks %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, r["avg_launch_ms"], {k:round(v,1) for k,v in r["stage_ms"].items()}, d["clocks"]["sm_mhz"]))
PY
done
