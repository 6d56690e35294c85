#!/bin/bash
# round 2, GPU call A: new backward kernels + regression of the existing suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_backward.py -m gpu -q -s --maxfail=30 2>&1 | tail -150 > gpurun_out/r2a_backward.log
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py tests/test_gpu_blocks.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2a_regress.log
tail -40 gpurun_out/r2a_backward.log
tail -15 gpurun_out/r2a_regress.log
