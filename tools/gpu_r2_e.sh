#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2e_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 1 0 > gpurun_out/r2e_timeline_c1.txt 2>&1
timeout 300 python tools/attn_timeline.py 2 0 > gpurun_out/r2e_timeline_c2.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 30 > gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_attn_variants.txt; tail -8 gpurun_out/r2e_timeline_c1.txt; tail -8 gpurun_out/r2e_timeline_c2.txt; tail -n 6 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_bench.json; tail -5 gpurun_out/r2e_bench.err
