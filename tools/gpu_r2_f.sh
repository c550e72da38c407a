#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2f_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 0 0 > gpurun_out/r2f_timeline.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_splat_launches.csv python tools/splat_bench.py --profile > gpurun_out/r2f_splat_ncu.log 2>&1
timeout 300 python tools/splat_bench.py > gpurun_out/r2f_splat.json 2>&1
timeout 900 python -m pytest tests/test_raster_gpu.py tests/test_renderer_gpu.py tests/test_attention_adversarial_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 30 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_attn_variants.txt; tail -8 gpurun_out/r2f_timeline.txt; cat gpurun_out/r2f_splat.json; tail -n 8 gpurun_out/r2f_pytest.log
