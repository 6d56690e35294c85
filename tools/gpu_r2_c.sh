#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_backward.py -m gpu -q -s -x -k "every_conv or finite" > gpurun_out/sanitize_pytest.log 2>&1
grep -E "Invalid|out of bounds|ERROR SUMMARY|passed|failed|Address|=========     at|by thread|RuntimeError" gpurun_out/sanitize_pytest.log | head -60
cat gpurun_out/fd_report.txt
