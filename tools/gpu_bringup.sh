#!/bin/bash
# Run the per-kernel GPU parity tests in isolated processes (a trapped kernel poisons its CUDA context) and collect logs.
# Usage (on the GPU box, through gpurun): bash tools/gpu_bringup.sh [extra pytest args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/bringup_gpu.txt 2>&1
LOG=gpurun_out/bringup.log
: > $LOG
run() {
  echo "=== $1" >> $LOG
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "$1" "${@:2}" 2>&1 | tail -n 60 >> $LOG
}
run "test_gemm and simt"
run "test_gemm and tc"
run "test_conv3x3 and simt"
run "test_conv3x3 and tc and s1"
run "test_conv3x3 and tc and s2"
run "test_cross_view_attention and simt"
run "test_cross_view_attention and tc and l0"
run "test_cross_view_attention and tc and l1"
run "test_cross_view_attention and tc and (l2 or l3)"
run "test_cross_attention_text_keys and simt"
run "test_cross_attention_text_keys and tc"
run "test_group_norm or test_layer_norm or test_temporal or test_conv_in or test_ddim"
grep -E "^===|passed|failed|error" $LOG | head -80
