#!/bin/bash
# 8 GPUs: (1) the NCCL worker of tests/test_multigpu_gpu.py at world 4 (one view per rank, checked against the unsharded forward),
# (2) bench at N = 8: prompt-sharded weak scaling + one prompt over 2 CFG x 4 view ranks
mkdir -p gpurun_out
python - <<'P'
import re, textwrap
src = open("tests/test_multigpu_gpu.py").read()
m = re.search(r"WORKER = textwrap\.dedent\(r'''(.*?)'''\)", src, re.S)
open("gpurun_out/mgpu_worker4.py", "w").write(textwrap.dedent(m.group(1)))
P
A3D_ROOT=$PWD timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 gpurun_out/mgpu_worker4.py > gpurun_out/r2t_worker4.log 2>&1
echo "worker rc=$?" >> gpurun_out/r2t_worker4.log
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2t_bench8.json 2> gpurun_out/r2t_bench8.err
echo "rc=$?" >> gpurun_out/r2t_bench8.err
grep "OK\|rc=\|Error\|error\|rel" gpurun_out/r2t_worker4.log | tail -n 12
grep "^{" gpurun_out/r2t_bench8.json | cut -c1-400; grep -o '"view_sharded": {[^}]*}' gpurun_out/r2t_bench8.json; grep -o '"fwd_mpix_s": [0-9.]*, "fwd_bwd_mpix_s": [0-9.]*' gpurun_out/r2t_bench8.json; tail -n 4 gpurun_out/r2t_bench8.err
