#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_timeline.py 0 > gpurun_out/r2c_timeline0.txt 2>&1
timeout 300 python tools/attn_timeline.py 2 > gpurun_out/r2c_timeline2.txt 2>&1
timeout 600 python -m pytest tests/test_renderer_gpu.py tests/test_attention_adversarial_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 30 > gpurun_out/r2c_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_splat_launches.csv python tools/splat_bench.py > gpurun_out/r2c_splat_ncu.log 2>&1
timeout 300 python tools/splat_bench.py > gpurun_out/r2c_splat.json 2>&1
cat gpurun_out/r2c_timeline0.txt; tail -12 gpurun_out/r2c_timeline2.txt; tail -n 12 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_splat.json
