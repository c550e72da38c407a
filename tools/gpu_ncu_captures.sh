#!/bin/bash
# ncu captures for profiles/: --set full of the dominant attention kernel (default mode) and of the two render kernels
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:attn5_tc -s 2 -c 1 -f -o gpurun_out/r2n_attn5 python tools/kernel_bench.py attn0 > gpurun_out/r2n_ncu_attn.log 2>&1
timeout 900 $NCU --profile-from-start off -k regex:'render|deform_backward|radix|preprocess' -c 14 -f -o gpurun_out/r2n_splat python tools/splat_bench.py --profile > gpurun_out/r2n_ncu_splat.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2n_ncu_attn.log gpurun_out/r2n_ncu_splat.log
