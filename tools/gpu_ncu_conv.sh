#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/conv32_full python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu full rc=$?"
