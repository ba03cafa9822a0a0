#!/bin/bash
# compute-sanitizer over the kernels added in round 2: gmm_stream_kernel, fe_spectrum_wfft, tc16_norm_replay, utt_checksum_kernel,
# the regression-class branch of gmm_diag_f64, fe_build_row_utt
mkdir -p gpurun_out
K='streaming_scorer_equals or floor or device_buffers_and_overflow or per_utterance or checksum_sink or fused or warp_fft or launch_shape or regression or all_pass'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 10 --log-file gpurun_out/r02b_sanitizer_$tool.log \
    python -m pytest tests/test_gpu_stream.py tests/test_gpu_multigpu.py tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "$K" > gpurun_out/r02b_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02b_sanitizer_${tool}_pytest.log
  tail -2 gpurun_out/r02b_sanitizer_$tool.log; tail -2 gpurun_out/r02b_sanitizer_${tool}_pytest.log
done
grep "Error: Race" gpurun_out/r02b_sanitizer_racecheck.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | cut -c1-220
