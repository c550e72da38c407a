#!/bin/bash
# round-2 final evidence on one B200: smoke, full GPU suite, the bench line (both arms), the ncu launch list of one eager step and a
# --set full capture of the dominant kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2w_smoke.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -n 15 > gpurun_out/r2w_pytest.log
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_launches.csv python tools/one_step.py > gpurun_out/r2w_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn5_tc -s 2 -c 1 -f -o gpurun_out/r2w_attn5 python tools/kernel_bench.py attn0 > gpurun_out/r2w_ncu_attn.log 2>&1
timeout 300 python tools/op_breakdown.py > gpurun_out/r2w_breakdown.txt 2>&1
tail -n 2 gpurun_out/r2w_smoke.log; tail -n 5 gpurun_out/r2w_pytest.log
python - <<'P'
import json
for f in ('gpurun_out/r2w_bench.json','gpurun_out/r2w_bench_ref.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print({k:d.get(k) for k in ('impl','value','ms_per_step','e2e','gpu_launches','clocks')}); 
            if 'roofline' in d: print(d['roofline']['ms_per_launch'], d['roofline']['frac'], d['roofline']['mufu'], d['roofline_gemm']['short_k']['frac'], d['roofline_gemm']['long_k']['frac'], d['splat']['fwd_mpix_s'], d['splat']['fwd_bwd_mpix_s'], d['sds_step'] and d['sds_step'].get('ms_per_step'))
P
tail -n 2 gpurun_out/r2w_bench.err gpurun_out/r2w_bench_ref.err; wc -l gpurun_out/r2w_launches.csv; head -12 gpurun_out/r2w_breakdown.txt
