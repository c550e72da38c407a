#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run9.log
: > $LOG
run() {
  echo "=== $1 ($2)" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 25 >> $LOG
}
run "test_gemm and tc" tests/test_kernels_gpu.py
run "test_conv3x3 and tc" tests/test_kernels_gpu.py
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
echo "=== attention trace" >> $LOG
timeout 300 python tools/attn_trace.py >> $LOG 2>&1
echo "=== kernel bench" >> $LOG
timeout 300 python tools/kernel_bench.py all >> $LOG 2>&1
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | tail -n 12 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01d.json 2>> $LOG
cat gpurun_out/bench_r01d.json >> $LOG
tail -n 70 $LOG
