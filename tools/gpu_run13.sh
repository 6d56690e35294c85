#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_all.log
BPB_CONV_ROLES=1 timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv_roles.log 2>&1; grep -E "^conv|span|mma0|epilogue0|producer" gpurun_out/probe_conv_roles.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench13.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench13.log | cut -c1-200
