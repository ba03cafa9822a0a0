#!/bin/bash
# N = 2: the config-4 path (model broadcast, frame-count all-gather, LPT, per-rank writers, both LNA gathers) next to the headline
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,P2P NCCL_DEBUG_FILE=gpurun_out/r02_nccl_n2.%h.%p.log
AKUGPU_BENCH_TRACE= timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --c4-utts ${C4_UTTS:-900} > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "rc=$?"
tail -c 3000 gpurun_out/r02_bench_n2.json
grep -v "NCCL INFO" gpurun_out/r02_bench_n2.err | tail -25
ls gpurun_out/r02_nccl_n2.* 2>/dev/null | head -3
for f in gpurun_out/r02_nccl_n2.*; do grep -cE "Broadcast|AllGather|Send|Recv" $f; done
