#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run8.log
: > $LOG
run() {
  echo "=== $1 ($2)" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 25 >> $LOG
}
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
echo "=== attention trace" >> $LOG
timeout 300 python tools/attn_trace.py >> $LOG 2>&1
echo "=== kernel bench attn" >> $LOG
timeout 300 python tools/kernel_bench.py attn >> $LOG 2>&1
tail -n 45 $LOG
