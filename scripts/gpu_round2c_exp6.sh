#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1_final3.json 2> gpurun_out/r02c_bench_n1_final3.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r02c_bench_n1_final3.json"))
print(d["value"], d["e2e"]["value"], d["sub_records"]["parity_mode_f64"].get("throughput_mode_codes_vs_parity_mode"), d["sub_records"]["parity_mode_f64"].get("failed"))
PY
