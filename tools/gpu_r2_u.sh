#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes > gpurun_out/r2u_pipes.txt 2>&1
timeout 600 python tools/attn_variants.py > gpurun_out/r2u_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 31 > gpurun_out/r2u_timeline.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r2u_pytest.log
grep "chained\|^MUFU.EX2\|^F2FP\|^FMNMX3\|^FFMA2" gpurun_out/r2u_pipes.txt; grep "^attn" gpurun_out/r2u_attn_variants.txt; grep "g=0" gpurun_out/r2u_timeline.txt | tail -n 4; tail -n 6 gpurun_out/r2u_timeline.txt; tail -n 3 gpurun_out/r2u_pytest.log
