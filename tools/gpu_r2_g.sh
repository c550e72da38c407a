#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED\|^$" | tail -n 120 > gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_splat_launches.csv python tools/splat_bench.py --profile > gpurun_out/r2g_splat_ncu.log 2>&1
tail -n 25 gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_bench.json; tail -3 gpurun_out/r2g_bench.err
