#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_vae_gpu.py tests/test_kernels_gpu.py tests/test_processors_gpu.py tests/test_deform.py -q -m gpu --tb=short -p no:cacheprovider -rA 2>&1 | grep -v "^PASSED\|^$" | tail -n 100 > gpurun_out/r2h_pytest.log
tail -n 60 gpurun_out/r2h_pytest.log
