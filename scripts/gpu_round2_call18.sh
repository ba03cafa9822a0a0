#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_gputest_18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_18.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_gputest_18.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmm_stream_kernel -s 4 -c 1 -o gpurun_out/r02_gmm_stream_final -f python scripts/ncu_stream.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmm_stream_kernel -s 4 -c 1 -o gpurun_out/r02_gmm_stream_c4 -f python scripts/ncu_stream.py 10000 32 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gmm_tc16_kernel -s 1 -c 1 -o gpurun_out/r02_gmm_tc16 -f python scripts/ncu_gmm.py 0 30 2>&1 | tail -1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench_200utts.csv python bench.py --steps 1 --warmup 1 --utts 200 --no-cpu-baseline --no-sub-records > /dev/null 2>&1; wc -l gpurun_out/r02_launches_bench_200utts.csv
