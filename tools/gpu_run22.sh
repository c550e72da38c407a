#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run22.log
: > $LOG
echo "=== temporal tests" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "temporal" 2>&1 | tail -n 8 >> $LOG
echo "=== pipes ubench" >> $LOG
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes 2>&1 | grep -E "f16x2|MUFU.EX2  " >> $LOG
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py 2>&1 | grep -E "total|gemm=|temporal" >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01m.json 2>> $LOG
cat gpurun_out/bench_r01m.json >> $LOG
tail -n 40 $LOG
