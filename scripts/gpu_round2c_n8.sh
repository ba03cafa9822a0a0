#!/bin/bash
# N = 8 (or N = $1) on the final code of the third session: the driver's own command line, with the full-size config-4 sub-record
N=${1:-8}
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
timeout 1300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err
echo "rc=$?"
grep -v "NCCL INFO" gpurun_out/r02c_bench_n$N.err | grep -v "^\*\|OMP_NUM" | tail -12
head -c 300 gpurun_out/r02c_bench_n$N.json
