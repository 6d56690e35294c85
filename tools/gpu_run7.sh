#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run7.log
: > $L
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > $L 2>&1; echo "pytest rc=$?"
tail -5 $L
timeout 300 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; tail -6 gpurun_out/probe_eval.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench7.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench7.log | cut -c1-300
