#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2s_attn_variants.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r2s_pytest.log
timeout 300 python tools/op_breakdown.py > gpurun_out/r2s_breakdown.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
grep "^attn" gpurun_out/r2s_attn_variants.txt; tail -n 4 gpurun_out/r2s_pytest.log; head -40 gpurun_out/r2s_breakdown.txt; python - <<'P'
import json
for l in open('gpurun_out/r2s_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['ms_per_launch'], d['roofline']['mufu'], d['clocks'], d['e2e'], json.dumps(d.get('sds_step'))[:400])
P
tail -n 3 gpurun_out/r2s_bench.err
