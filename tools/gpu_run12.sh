#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run12.log
: > $LOG
run() {
  echo "=== $1 ($2) mode=$A3D_ATTN_MODE" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 15 >> $LOG
}
for m in 8 9; do
  export A3D_ATTN_MODE=$m
  run "attention and tc" tests/test_kernels_gpu.py
done
unset A3D_ATTN_MODE
for m in 0 1 8 9 10; do
  echo "=== kernel bench attn mode $m" >> $LOG
  A3D_ATTN_MODE=$m timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
done
for m in 8 9; do
echo "=== attention trace mode $m" >> $LOG
A3D_ATTN_MODE=$m timeout 300 python tools/attn_trace.py 2>&1 | tail -n 20 >> $LOG
done
for m in 0 8 9; do
echo "=== unet parity mode $m" >> $LOG
A3D_ATTN_MODE=$m timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -E "rel|passed|failed" | tail -n 8 >> $LOG
done
tail -n 120 $LOG
