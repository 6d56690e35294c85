#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "distance or full_size or medium" > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_dist.log
timeout 40 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; head -2 gpurun_out/probe_eval.log
