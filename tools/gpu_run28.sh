#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run28.log
: > $LOG
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude"
for st in 2 3 4; do
  echo "=== stages $st" >> $LOG
  $NV -DA3D_STAGES256=$st -c animate3d_b200/csrc/a3d_gemm.cu -o /tmp/a3d_gemm_$st.o 2>> $LOG
  $NV -shared -o animate3d_b200/liba3d.so /tmp/a3d_gemm_$st.o $(ls animate3d_b200/build/*.o | grep -v a3d_gemm) -gencode arch=compute_100a,code=sm_100a 2>> $LOG
  timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "qkv|proj |geglu|ffout|l0 conv" >> $LOG
done
tail -n 40 $LOG
