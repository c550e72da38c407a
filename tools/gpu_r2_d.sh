#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_timeline.py 0 > gpurun_out/r2d_timeline0.txt 2>&1
timeout 600 python tools/attn_variants.py > gpurun_out/r2d_attn_variants.txt 2>&1
timeout 900 python -m pytest tests/test_renderer_gpu.py tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 30 > gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_timeline0.txt; cat gpurun_out/r2d_attn_variants.txt; tail -n 12 gpurun_out/r2d_pytest.log
