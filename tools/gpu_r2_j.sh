#!/bin/bash
# 2 GPUs, tight timeouts: the sharded worker in eager mode, then with CUDA-graph capture of the NCCL gathers
mkdir -p gpurun_out
export A3D_ROOT=$PWD
A3D_SHARDED_GRAPH=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/mgpu_worker.py > gpurun_out/r2j_eager.log 2>&1
echo "eager rc=$?" >> gpurun_out/r2j_eager.log
A3D_SHARDED_GRAPH=1 NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/mgpu_worker.py > gpurun_out/r2j_graph.log 2>&1
echo "graph rc=$?" >> gpurun_out/r2j_graph.log
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2j_eager.log | tail -n 15; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2j_graph.log | tail -n 25
