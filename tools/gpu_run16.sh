#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -x -q -m gpu -k "distance or rank or simt or eval or full" > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dist.log
timeout 300 python tools/probe_eval.py > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; tail -6 gpurun_out/probe_eval.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench16.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/conv32_full_r1 python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_distance_tc -s 1 -c 1 -o gpurun_out/dist_full_r1 python tools/probe_eval.py > gpurun_out/ncu_dist.log 2>&1; echo "ncu dist rc=$?"
