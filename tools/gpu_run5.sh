#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run5.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python -m pytest -q --tb=short -p no:cacheprovider "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tests/test_gpu_kernels.py -x
run tests/test_gpu_path.py
run tests/test_gpu_blocks.py -s
grep -E "^(===|rc=|FAILED|ERROR)|passed|failed|blocks checked" $L | tail -40
grep -E "^E  " $L | cut -c1-300 | head -30
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; tail -14 gpurun_out/probe_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5_graph.log 2>&1; echo "bench graph rc=$?"; tail -1 gpurun_out/bench5_graph.log | cut -c1-2500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/conv32_full python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu full rc=$?"
