#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run3.log
: > $LOG
run() {
  echo "=== $1" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 40 >> $LOG
}
run "test_gemm and tc" tests/test_kernels_gpu.py
run "test_conv3x3 and tc" tests/test_kernels_gpu.py
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
run "indices" tests/test_raster_gpu.py
run "backward" tests/test_raster_gpu.py
run "reference_style" tests/test_raster_gpu.py
echo "=== kernel bench (v2 attention)" >> $LOG
timeout 300 python tools/kernel_bench.py all >> $LOG 2>&1
echo "=== kernel bench attention v1" >> $LOG
A3D_ATTN_VARIANT=1 timeout 300 python tools/kernel_bench.py attn >> $LOG 2>&1
echo "=== ncu attention v2" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_tc_kernel -s 3 -c 1 -o gpurun_out/prof_attn2 python tools/kernel_bench.py attn >> $LOG 2>&1
echo "=== ncu gemm qkv" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 42 -c 1 -o gpurun_out/prof_gemm_qkv python tools/kernel_bench.py gemm >> $LOG 2>&1
tail -n 120 $LOG
