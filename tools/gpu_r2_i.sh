#!/bin/bash
# 2-GPU validation: NCCL tests (view / CFG / camera sharding) + bench under torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2i_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_vae_gpu.py tests/test_raster_gpu.py tests/test_renderer_gpu.py tests/test_deform.py -q -m gpu --tb=short -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED\|^$" | tail -n 60 > gpurun_out/r2i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err
tail -n 30 gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_bench2.json; tail -n 15 gpurun_out/r2i_bench2.err
