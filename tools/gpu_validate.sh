#!/bin/bash
# Full GPU validation on one B200 (run through gpurun): parity tests, smoke, the bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log
