#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_k.log | cut -c1-200
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; tail -12 gpurun_out/probe_conv.log
for i in 1 2; do timeout 900 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench20_$i.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench20_$i.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['ms_per_launch'])"; done
