#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/e2e_report.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -60 > gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; tail -8 gpurun_out/r2d_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 3000 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
