#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2b_attn_variants.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_renderer_gpu.py tests/test_raster_gpu.py tests/test_arap_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 40 > gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_attn_variants.txt; tail -n 25 gpurun_out/r2b_pytest.log
