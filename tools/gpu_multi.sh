#!/bin/bash
# run with: gpurun --gpus 2 -- 'bash tools/gpu_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out
LOG=gpurun_out/multi_$N.log
: > $LOG
nvidia-smi --query-gpu=index,name --format=csv >> $LOG 2>&1
echo "=== bench --gpus $N (prompt-sharded, torchrun)" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_multi_$N.json 2>> $LOG
cat gpurun_out/bench_multi_$N.json >> $LOG
echo "=== view-parallel check over $N ranks" >> $LOG
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/view_parallel_check.py 4 >> $LOG 2>&1
tail -n 40 $LOG
