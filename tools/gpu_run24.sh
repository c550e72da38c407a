#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run24.log
: > $LOG
echo "=== attention tests mode 3" >> $LOG
A3D_ATTN_MODE=3 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -n 8 >> $LOG
for m in 0 3 0 3; do
echo "=== kernel bench attn v5 mode $m" >> $LOG
A3D_ATTN_MODE=$m timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
done
echo "=== unet parity mode 3" >> $LOG
A3D_ATTN_MODE=3 timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
tail -n 40 $LOG
