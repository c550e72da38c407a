#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run32.log
: > $LOG
echo "=== gemm/conv tests" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x -k "gemm or conv3x3" 2>&1 | tail -n 6 >> $LOG
echo "=== kernel bench gemm" >> $LOG
timeout 300 python tools/kernel_bench.py gemm >> $LOG 2>&1
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py 2>&1 | grep -E "total|gemm=" >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01p.json 2>> $LOG
cat gpurun_out/bench_r01p.json >> $LOG
tail -n 60 $LOG
