#!/bin/bash
# A/B of the LayerNorm-into-GEMM fold: meaningful only with profiles/r02_ln_fold_experiment.patch applied (A3D_FUSE_LN is read by that patch).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "gemm or layer_norm" 2>&1 | tail -n 15 > gpurun_out/r2x1_pytest.log
A3D_FUSE_LN=0 timeout 300 python tools/op_breakdown.py > gpurun_out/r2x1_breakdown_unfused.txt 2>&1
A3D_FUSE_LN=1 timeout 300 python tools/op_breakdown.py > gpurun_out/r2x1_breakdown_fused.txt 2>&1
timeout 600 python -m pytest tests/test_unet_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "plumbing or cfg or multi" 2>&1 | tail -n 8 > gpurun_out/r2x1_pytest_unet.log
tail -n 6 gpurun_out/r2x1_pytest.log; tail -n 5 gpurun_out/r2x1_pytest_unet.log
for f in unfused fused; do echo "== $f"; sed -n 2,4p gpurun_out/r2x1_breakdown_$f.txt; grep "layer_norm\| LN\|N=2560 K=320\|N=1152 K=320\|N=960 K=320\|N=1536 K=320" gpurun_out/r2x1_breakdown_$f.txt | head -12; done
