#!/bin/bash
# launch list of the final code (bench.py --steps 1 --warmup 1 --utts 200) + ncu --set full of the scorer in its final state
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_bench_200utts.csv python bench.py --steps 1 --warmup 1 --utts 200 --no-cpu-baseline --no-sub-records > /dev/null 2>&1; wc -l gpurun_out/r02c_launches_bench_200utts.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gmm_tc16_kernel -s 1 -c 1 -o gpurun_out/r02c_gmm_tc16 -f python scripts/ncu_gmm.py 0 30 2>&1 | tail -1
