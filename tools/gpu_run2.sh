#!/bin/bash
# second GPU run: full test-suite in isolated processes + bench (graph / no graph) + ncu launch list of one step
mkdir -p gpurun_out
L=gpurun_out/run2.log
: > $L
run() { echo "=== $*" >> $L; timeout 600 python -m pytest -q --tb=short -p no:cacheprovider "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tests/test_gpu_kernels.py -k "conv_tc and 0-"     # placeholder, replaced below
: > $L
run tests/test_gpu_kernels.py -k "linear"
run tests/test_gpu_kernels.py -k "conv"
run tests/test_gpu_kernels.py -k "not conv and not linear"
run tests/test_gpu_path.py -k "gilt or triplet or body_part or distance or signature or rank"
run tests/test_gpu_path.py -k "medium_scale or full_size"
run tests/test_gpu_path.py -k "model or conv_paths or training_forward or map_parity"
run tests/test_gpu_blocks.py -s
grep -E "^(===|rc=|FAILED|ERROR)|passed|failed|blocks checked" $L | tail -60
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --skip-cpu > gpurun_out/bench2_nograph.log 2>&1; echo "bench nograph rc=$?"; tail -1 gpurun_out/bench2_nograph.log | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2_graph.log 2>&1; echo "bench graph rc=$?"; tail -1 gpurun_out/bench2_graph.log | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md | head -40
