#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2k_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 3 > gpurun_out/r2k_timeline_ts.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_vae_gpu.py tests/test_raster_gpu.py tests/test_renderer_gpu.py tests/test_deform.py -q -m gpu --tb=short -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED\|^$" | tail -n 60 > gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
timeout 300 python tools/op_breakdown.py > gpurun_out/r2k_breakdown.txt 2>&1
cat gpurun_out/r2k_attn_variants.txt; tail -8 gpurun_out/r2k_timeline_ts.txt; tail -n 14 gpurun_out/r2k_pytest.log; cut -c1-1500 gpurun_out/r2k_bench.json; head -30 gpurun_out/r2k_breakdown.txt
