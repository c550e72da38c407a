#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/attn_variants.py > gpurun_out/r2p_attn_variants.txt 2>&1
timeout 300 python tools/attn_timeline.py 7 > gpurun_out/r2p_timeline.txt 2>&1
timeout 900 python -m pytest tests/test_attention_adversarial_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 12 > gpurun_out/r2p_pytest.log
timeout 300 python tools/op_breakdown.py > gpurun_out/r2p_breakdown.txt 2>&1
timeout 300 python tools/kernel_bench.py gemm > gpurun_out/r2p_kernel_bench.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
cat gpurun_out/r2p_attn_variants.txt; tail -14 gpurun_out/r2p_timeline.txt; tail -n 6 gpurun_out/r2p_pytest.log; head -24 gpurun_out/r2p_breakdown.txt; cat gpurun_out/r2p_kernel_bench.txt; python - <<'P'
import json
for l in open('gpurun_out/r2p_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['ms_per_launch'], d['roofline']['mufu'])
P
tail -n 3 gpurun_out/r2p_bench.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn5_tc -s 2 -c 1 -f -o gpurun_out/r2n_attn5 python tools/kernel_bench.py attn0 > gpurun_out/r2n_ncu_attn.log 2>&1
tail -n 3 gpurun_out/r2n_ncu_attn.log
