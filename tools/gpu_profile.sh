#!/bin/bash
# Profiles for profiles/: ncu launch lists of one training step and of the evaluation probe, full captures of the two
# dominant kernels (never a timing source: bench.py / probe_*.py numbers come from un-profiled runs).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --profile > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python tools/summarize_launches.py gpurun_out/launches_r1.csv gpurun_out/launches_r1.md > /dev/null 2>&1; head -16 gpurun_out/launches_r1.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_eval.csv python tools/probe_eval.py > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/conv32_full_r1 python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_distance_tc -s 1 -c 1 -o gpurun_out/dist_full_r1 python tools/probe_eval.py > gpurun_out/ncu_dist.log 2>&1; echo "ncu dist rc=$?"
