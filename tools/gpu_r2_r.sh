#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2r_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 19 > gpurun_out/r2r_timeline.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r2r_pytest.log
cat gpurun_out/r2r_attn_variants.txt; grep "step 2[01]\|step 3[01]" gpurun_out/r2r_timeline.txt; tail -n 10 gpurun_out/r2r_timeline.txt;  tail -n 4 gpurun_out/r2r_pytest.log
