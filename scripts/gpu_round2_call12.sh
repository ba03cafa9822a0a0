#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_refbin.py -m gpu -q -s -k "feature or fused or warp_fft or sweep or refbin or goldens or vtln or speaker or pre" > gpurun_out/r02_gputest_12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_12.log
grep -E "passed|failed|FAILED|ERROR|rc=" gpurun_out/r02_gputest_12.log | tail -8
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_new.json 2> gpurun_out/r02_bench_c3_new.err; grep "config 3" gpurun_out/r02_bench_c3_new.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_spectrum_wfft -s 2 -c 1 -o gpurun_out/r02_fe_wfft -f python scripts/ncu_fe.py 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fe_delta2_merge -s 2 -c 1 -o gpurun_out/r02_fe_delta2 -f python scripts/ncu_fe.py 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gmm_stream_kernel -s 4 -c 1 -o gpurun_out/r02_gmm_stream -f python scripts/ncu_stream.py 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
