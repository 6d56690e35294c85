#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
bash tools/gpu_r2_e.sh > gpurun_out/r2j_launches.log 2>&1; head -30 gpurun_out/launches_r2_train.md
bash tools/gpu_r2_profile.sh > gpurun_out/r2j_profile.log 2>&1; tail -30 gpurun_out/r2j_profile.log
