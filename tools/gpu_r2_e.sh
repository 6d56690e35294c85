#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_train.csv python bench.py --workload train --steps 1 --warmup 3 --no-graph --skip-cpu --no-extras --profile > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
python tools/summarize_step.py gpurun_out/launches_r2_train.csv gpurun_out/launches_r2_train.md "one training step (fwd + loss + backward), HRNet-W32 bs256 256x128" | head -50
