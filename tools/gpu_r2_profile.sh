#!/bin/bash
# round 2 profiles: full ncu captures of the three dominant kernels + per-shape timings (never a bench value)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/probe_conv.py all > gpurun_out/conv_probe_r2.txt 2>&1; python tools/probe_wgrad.py all > gpurun_out/wgrad_probe_r2.txt 2>&1
cat gpurun_out/conv_probe_r2.txt gpurun_out/wgrad_probe_r2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 2 -c 1 -f -o gpurun_out/conv32_full_r2 python tools/probe_conv.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_stacked_kernel -s 2 -c 1 -f -o gpurun_out/wgrad32_stacked_full_r2 python tools/probe_wgrad.py > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:part_distance_tc -s 1 -c 1 -f -o gpurun_out/dist_full_r2 python tools/probe_eval.py > gpurun_out/ncu_dist.log 2>&1; echo "ncu dist rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce -s 40 -c 1 -f -o gpurun_out/bnred_full_r2 python bench.py --workload train --steps 1 --warmup 3 --no-graph --skip-cpu --no-extras --profile > gpurun_out/ncu_bnred.log 2>&1; echo "ncu bnred rc=$?"
python tools/probe_bnbwd.py > gpurun_out/bnbwd_probe_r2.txt 2>&1; cat gpurun_out/bnbwd_probe_r2.txt
ls -la gpurun_out/*_r2.ncu-rep
