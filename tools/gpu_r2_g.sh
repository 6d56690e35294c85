#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_dist_train.py > gpurun_out/r2g_dist_train.log 2>&1; tail -5 gpurun_out/r2g_dist_train.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 10 --warmup 3 --skip-cpu --no-extras > gpurun_out/r2g_n2_train.json 2> gpurun_out/r2g_n2_train.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_n2_train.json')); print('train N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/r2g_n2_train.err
