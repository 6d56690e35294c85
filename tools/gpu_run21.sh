#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_p.log
timeout 900 python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench21.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench21.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['ms_per_launch'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md > /dev/null 2>&1; head -22 gpurun_out/launches_r1.md
