#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/umma_rate > gpurun_out/umma_rate2.log 2>&1; cat gpurun_out/umma_rate2.log
BPB_CONV_ROLES=1 timeout 300 python tools/probe_conv.py all > gpurun_out/probe_conv_roles.log 2>&1; grep -E "^conv|CTAs" gpurun_out/probe_conv_roles.log
