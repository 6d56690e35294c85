#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_lanes.py > gpurun_out/diag_lanes.log 2>&1; echo "diag rc=$?"; tail -14 gpurun_out/diag_lanes.log
