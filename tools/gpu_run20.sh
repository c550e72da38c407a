#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run20.log
: > $LOG
echo "=== gemm/conv/gn tests" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "gemm or conv3x3 or group_norm" 2>&1 | tail -n 15 >> $LOG
echo "=== kernel bench gemm (heuristic)" >> $LOG
timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj " >> $LOG
for bn in 128 160 256; do
echo "=== kernel bench gemm forced BN=$bn" >> $LOG
A3D_GEMM_BN=$bn timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj |ffout" >> $LOG
done
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed|rror" | tail -n 8 >> $LOG
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py 2>&1 | head -70 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01k.json 2>> $LOG
cat gpurun_out/bench_r01k.json >> $LOG
tail -n 130 $LOG
