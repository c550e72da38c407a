#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run33.log
: > $LOG
echo "=== conv_in/out tests" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "conv_in_out" 2>&1 | tail -n 8 >> $LOG
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py 2>&1 | grep -E "total|gemm=|conv_in|conv_out|upsample" >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01q.json 2>> $LOG
cat gpurun_out/bench_r01q.json >> $LOG
tail -n 40 $LOG
