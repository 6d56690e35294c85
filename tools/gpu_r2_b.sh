#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -s -k "not conv_wgrad_tc and not bn_backward_kernels" 2>&1 | tail -150 > gpurun_out/r2b_backward.log
tail -100 gpurun_out/r2b_backward.log
