#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run31.log
: > $LOG
echo "=== gemm/conv tests (tma store)" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x -k "gemm or conv3x3" 2>&1 | tail -n 15 >> $LOG
echo "=== kernel bench gemm tma store" >> $LOG
timeout 300 python tools/kernel_bench.py gemm >> $LOG 2>&1
echo "=== kernel bench gemm direct stores" >> $LOG
A3D_GEMM_TMA_STORE=0 timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj|ffout|geglu" >> $LOG
echo "=== kernel bench N=320 with BN=128 (tma)" >> $LOG
A3D_GEMM_BN=128 timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "proj|ffout|l0 conv" >> $LOG
tail -n 60 $LOG
