#!/bin/bash
# round-2 first GPU pass: whole GPU suite (no -x: list every failure), bench line, op breakdown
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED\|^$" | tail -n 150 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 300 python tools/op_breakdown.py > gpurun_out/r2a_breakdown.txt 2>&1
tail -n 30 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_bench.json; head -n 40 gpurun_out/r2a_breakdown.txt
