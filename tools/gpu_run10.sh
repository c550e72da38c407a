#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run10.log
: > $LOG
run() {
  echo "=== $1 ($2) mode=$A3D_ATTN_MODE" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 15 >> $LOG
}
for m in 2 3 6; do
  export A3D_ATTN_MODE=$m
  run "attention and tc" tests/test_kernels_gpu.py
done
unset A3D_ATTN_MODE
for m in 0 1 2 3 4 6; do
  echo "=== kernel bench attn mode $m" >> $LOG
  A3D_ATTN_MODE=$m timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
done
echo "=== kernel bench attn variant 2" >> $LOG
A3D_ATTN_VARIANT=2 timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== attention trace (default mode)" >> $LOG
timeout 300 python tools/attn_trace.py 2>&1 | tail -n 30 >> $LOG
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | tail -n 12 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2>> $LOG
cat gpurun_out/bench_r01e.json >> $LOG
tail -n 120 $LOG
