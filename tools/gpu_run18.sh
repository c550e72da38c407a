#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run18.log
: > $LOG
echo "=== attention tests (v5)" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -n 25 >> $LOG
echo "=== kernel bench attn v5" >> $LOG
timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== kernel bench attn v5 poly1" >> $LOG
A3D_ATTN_MODE=1 timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== kernel bench attn v4" >> $LOG
A3D_ATTN_VARIANT=4 timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01i.json 2>> $LOG
cat gpurun_out/bench_r01i.json >> $LOG
tail -n 60 $LOG
