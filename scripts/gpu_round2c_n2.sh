#!/bin/bash
# N = 2 on the final code of the third session: the driver's command line (config-2 weak scaling + config-4 sub-record) and the
# two-process GPU tests
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err
echo "rc=$?"
grep -v "NCCL INFO" gpurun_out/r02c_bench_n2.err | tail -8
