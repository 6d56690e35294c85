#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_distance_tc -s 1 -c 1 -o gpurun_out/dist_full python tools/probe_eval.py > gpurun_out/ncu_dist.log 2>&1; echo "ncu rc=$?"
