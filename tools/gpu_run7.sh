#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run7.log
: > $LOG
run() {
  echo "=== $1 ($2)" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 25 >> $LOG
}
run "test_gemm and tc" tests/test_kernels_gpu.py
run "test_conv3x3 and tc" tests/test_kernels_gpu.py
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
echo "=== kernel bench" >> $LOG
timeout 300 python tools/kernel_bench.py all >> $LOG 2>&1
echo "=== kernel bench attn v2" >> $LOG
A3D_ATTN_VARIANT=2 timeout 300 python tools/kernel_bench.py attn >> $LOG 2>&1
tail -n 60 $LOG
echo "=== attention trace" >> $LOG
timeout 300 python tools/attn_trace.py >> $LOG 2>&1
tail -n 30 $LOG
