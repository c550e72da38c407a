#!/bin/bash
# 4 GPUs: bench with the view-sharded strong-scaling mode (tight timeout; the bench's own watchdog guards the sharded section)
mkdir -p gpurun_out
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2l_bench4.json 2> gpurun_out/r2l_bench4.err
echo "rc=$?" >> gpurun_out/r2l_bench4.err
grep "^{" gpurun_out/r2l_bench4.json | cut -c1-600; grep -o '"view_sharded": {[^}]*}' gpurun_out/r2l_bench4.json; grep -o '"fwd_mpix_s": [0-9.]*, "fwd_bwd_mpix_s": [0-9.]*' gpurun_out/r2l_bench4.json; tail -n 8 gpurun_out/r2l_bench4.err
