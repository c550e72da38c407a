#!/bin/bash
# 4 GPUs: (1) the NCCL worker of tests/test_multigpu_gpu.py at world 4 (4 views, one per rank, checked against the single-GPU forward;
# camera-sharded rasterizer), (2) bench with the view-sharded strong-scaling mode (the bench's own watchdog guards that section)
mkdir -p gpurun_out
python - <<'P'
import re
src = open("tests/test_multigpu_gpu.py").read()
m = re.search(r"WORKER = r?'''(.*?)'''", src, re.S)
open("gpurun_out/mgpu_worker4.py", "w").write(m.group(1))
P
A3D_ROOT=$PWD timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 gpurun_out/mgpu_worker4.py > gpurun_out/r2l_worker4.log 2>&1
echo "worker rc=$?" >> gpurun_out/r2l_worker4.log
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2l_bench4.json 2> gpurun_out/r2l_bench4.err
echo "rc=$?" >> gpurun_out/r2l_bench4.err
grep "OK\|rc=\|Error\|error" gpurun_out/r2l_worker4.log | tail -n 12
grep "^{" gpurun_out/r2l_bench4.json | cut -c1-600; grep -o '"view_sharded": {[^}]*}' gpurun_out/r2l_bench4.json; grep -o '"fwd_mpix_s": [0-9.]*, "fwd_bwd_mpix_s": [0-9.]*' gpurun_out/r2l_bench4.json; tail -n 8 gpurun_out/r2l_bench4.err
