#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run5.log
: > $LOG
run() {
  echo "=== $1" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 30 >> $LOG
}
run "test_gemm and tc" tests/test_kernels_gpu.py
run "test_conv3x3 and tc" tests/test_kernels_gpu.py
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
echo "=== kernel bench" >> $LOG
timeout 300 python tools/kernel_bench.py all >> $LOG 2>&1
echo "=== ncu attention v3" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn3 python tools/kernel_bench.py attn >> $LOG 2>&1
echo "=== ncu gemm qkv (epilogue A)" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 42 -c 1 -o gpurun_out/prof_gemm_qkv2 python tools/kernel_bench.py gemm >> $LOG 2>&1
tail -n 80 $LOG
