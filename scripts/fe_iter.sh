#!/bin/bash
# front-end iteration: config-3 bench line + instruction count / issue utilisation of the warp FFT kernel
mkdir -p gpurun_out
timeout 200 python bench.py --config 3 > gpurun_out/fe_iter_bench.json 2> gpurun_out/fe_iter_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/fe_iter_bench.json"))
print("value %.1f M frames/s" % (d["value"] / 1e6), [(s["sample_rate"], s["window"], round(s["frames_per_s"] / 1e6, 1)) for s in d["sweep"]])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:fe_spectrum_.fft -s 2 -c 1 --csv --log-file gpurun_out/fe_iter_ncu.csv python scripts/ncu_fe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/fe_iter_ncu.csv")) if len(r) > 5]
hdr = rows[0]
for r in rows[1:]:
    d = dict(zip(hdr, r))
    print(d.get("Metric Name"), d.get("Metric Value"), d.get("Metric Unit"))
PY
