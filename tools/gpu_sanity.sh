#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -n 2
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_attention_adversarial_gpu.py tests/test_raster_gpu.py tests/test_renderer_gpu.py tests/test_processors_gpu.py tests/test_arap_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 3
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'P'
import json
for l in open('gpurun_out/r2z_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['ms_per_launch'])
P
