#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run6.log
: > $LOG
run() {
  echo "=== $1 ($2)" >> $LOG
  timeout 600 python -m pytest $2 -q -m gpu --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -n 25 >> $LOG
}
run "test_cross_view_attention and tc" tests/test_kernels_gpu.py
run "test_cross_attention_text_keys and tc" tests/test_kernels_gpu.py
run "deform" tests/test_deform.py
run "batch_forward" tests/test_renderer_gpu.py
echo "=== kernel bench attn (v3 per-tile MMA warps)" >> $LOG
timeout 300 python tools/kernel_bench.py attn >> $LOG 2>&1
echo "=== splat bench" >> $LOG
timeout 600 python tools/splat_bench.py >> $LOG 2>&1
echo "=== unet parity" >> $LOG
timeout 900 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | tail -n 12 >> $LOG
echo "=== bench" >> $LOG
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2>> $LOG
cat gpurun_out/bench_r01c.json >> $LOG
echo "=== ncu launch list" >> $LOG
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv python tools/one_step.py >> $LOG 2>&1
tail -n 120 $LOG
