#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/tma_rate > gpurun_out/tma_rate.log 2>&1; echo "tma rc=$?"
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "distance or rank or simt or eval or full" > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_dist.log
timeout 300 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/probe_eval.log
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; tail -12 gpurun_out/probe_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench9.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench9.log | cut -c1-200
