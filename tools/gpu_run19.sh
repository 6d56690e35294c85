#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/library_baseline.py > gpurun_out/library_baseline.log 2>&1; echo "lib rc=$?"; tail -2 gpurun_out/library_baseline.log | cut -c1-600
