#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run16.log
: > $LOG
echo "=== attention tests" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -n 25 >> $LOG
for m in 0 1 2; do
echo "=== kernel bench attn v4 poly $m" >> $LOG
A3D_ATTN_MODE=$m timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
done
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py 2>&1 | head -40 >> $LOG
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01h.json 2>> $LOG
cat gpurun_out/bench_r01h.json >> $LOG
tail -n 90 $LOG
