#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "distance or rank or simt or eval or full" > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dist.log
timeout 300 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/probe_eval.log
timeout 900 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench17.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench17.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')})"
