#!/bin/bash
# final check after the LNA kernel change: the whole GPU suite, an ncu capture of the new kernel, the driver's N = 1 command line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_gputest_final2.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/r02c_gputest_final2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lna_f32_rows_tma -s 1 -c 1 -o gpurun_out/r02c_lna_tma -f python scripts/ncu_gmm.py 0 30 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_n1_final2.json 2> gpurun_out/r02c_bench_n1_final2.err; echo "bench rc=$?"
head -c 300 gpurun_out/r02c_bench_n1_final2.json
