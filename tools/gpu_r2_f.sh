#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "multi_lane" 2>&1 | tail -12
cat gpurun_out/lanes_report.txt
