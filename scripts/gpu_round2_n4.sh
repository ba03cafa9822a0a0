#!/bin/bash
# N = 8: the driver's own command line, with the full-size config-4 sub-record
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,P2P NCCL_DEBUG_FILE=gpurun_out/r02_nccl_n4.%h.%p.log
nvidia-smi topo -m > gpurun_out/r02_topo_n4.txt 2>&1
timeout 1300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus 4 --steps ${STEPS:-5} --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
echo "rc=$?"
grep -v "NCCL INFO" gpurun_out/r02_bench_n4.err | grep -v "^\*\|OMP_NUM" | tail -25
# keep one rank's NCCL log (small), drop the rest
ls gpurun_out/r02_nccl_n4.* | tail -n +2 | xargs rm -f
for f in gpurun_out/r02_nccl_n4.*; do grep -E "Broadcast|AllGather|Send|Recv" $f | awk '{print $5}' | sort | uniq -c; done
