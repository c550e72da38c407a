#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run14.log
: > $LOG
for sg in 0 400 700 900 1100 1400; do
  echo "=== kernel bench attn stagger $sg" >> $LOG
  A3D_ATTN_STAGGER=$sg timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
done
echo "=== kernel bench attn stagger 900 mode 1" >> $LOG
A3D_ATTN_STAGGER=900 A3D_ATTN_MODE=1 timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== kernel bench attn stagger 900 mode 2" >> $LOG
A3D_ATTN_STAGGER=900 A3D_ATTN_MODE=2 timeout 300 python tools/kernel_bench.py attn0 >> $LOG 2>&1
echo "=== attention trace stagger 900" >> $LOG
A3D_ATTN_STAGGER=900 timeout 300 python tools/attn_trace.py 2>&1 | tail -n 20 >> $LOG
export A3D_ATTN_STAGGER=900
echo "=== attention tests stagger 900" >> $LOG
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "attention and tc" 2>&1 | tail -n 5 >> $LOG
unset A3D_ATTN_STAGGER
echo "=== op breakdown (2nd step)" >> $LOG
timeout 600 python tools/op_breakdown.py >> $LOG 2>&1
tail -n 130 $LOG
