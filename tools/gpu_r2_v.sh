#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2v_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 15 > gpurun_out/r2v_timeline.txt 2>&1
timeout 1200 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r2v_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
grep "^attn" gpurun_out/r2v_attn_variants.txt; grep "g=0" gpurun_out/r2v_timeline.txt | tail -n 4; tail -n 6 gpurun_out/r2v_timeline.txt; tail -n 3 gpurun_out/r2v_pytest.log
python - <<'P'
import json
for l in open('gpurun_out/r2v_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['ms_per_launch'], d['roofline']['mufu'], d['clocks'])
P
tail -n 3 gpurun_out/r2v_bench.err
