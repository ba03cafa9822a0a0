#!/bin/bash
# ncu --set full of the front-end kernels of the third session + a config-3 bench line
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_spectrum_bfft -s 2 -c 1 -f -o gpurun_out/r02c_fe_bfft python scripts/ncu_fe.py > /dev/null 2>&1; echo "bfft rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_delta2_merge_tiled -s 2 -c 1 -f -o gpurun_out/r02c_fe_delta_tiled python scripts/ncu_fe.py > /dev/null 2>&1; echo "delta rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_fe.csv python scripts/ncu_fe.py > /dev/null 2>&1; echo "launches rc=$?"
timeout 200 python bench.py --config 3 > gpurun_out/r02c_bench_config3.json 2> gpurun_out/r02c_bench_config3.err; echo "bench rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
