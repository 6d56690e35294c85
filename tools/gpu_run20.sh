#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "distance or rank or simt or eval or full" > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_dist.log
timeout 300 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/probe_eval.log
./tools/gpu_profile.sh
