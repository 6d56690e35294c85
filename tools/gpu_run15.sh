#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_path.py tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/pytest_ms.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ms.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench15_ms.log 2>&1; echo "bench multi-stream rc=$?"; tail -1 gpurun_out/bench15_ms.log | cut -c1-220
BPB_SINGLE_STREAM=1 timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench15_ss.log 2>&1; echo "bench single-stream rc=$?"; tail -1 gpurun_out/bench15_ss.log | cut -c1-220
