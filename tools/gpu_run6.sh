#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run6.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python -m pytest -q --tb=short -p no:cacheprovider "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tests/test_gpu_kernels.py -x -k "conv or linear"
run tests/test_gpu_blocks.py -s
grep -E "^(===|rc=|FAILED|ERROR)|passed|failed|blocks checked" $L | tail -20
grep -E "^E  " $L | cut -c1-300 | head -10
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; tail -14 gpurun_out/probe_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench6_graph.log 2>&1; echo "bench graph rc=$?"; tail -1 gpurun_out/bench6_graph.log | cut -c1-400
