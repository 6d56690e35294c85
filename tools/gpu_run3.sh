#!/bin/bash
# third GPU run: tests, bench, launch list (bounded), one ncu --set full capture of the dominant conv kernel
mkdir -p gpurun_out
L=gpurun_out/run3.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python -m pytest -q --tb=short -p no:cacheprovider "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tests/test_gpu_kernels.py
run tests/test_gpu_path.py -k "not full_size"
run tests/test_gpu_path.py -k "full_size"
run tests/test_gpu_blocks.py -s
grep -E "^(===|rc=|FAILED|ERROR)|passed|failed|blocks checked" $L | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3_graph.log 2>&1; echo "bench graph rc=$?"; tail -1 gpurun_out/bench3_graph.log | cut -c1-1800
# launch list of ONE step (skip the 3 warm-up steps' launches: 4 x ~700)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 900 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md | head -40
# full capture of the dominant conv kernel (probe script: a few launches of the C=32 3x3 conv at bs256)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 2 -o gpurun_out/conv32_full python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv.log 2>&1; cat gpurun_out/probe_conv.log | tail -30
