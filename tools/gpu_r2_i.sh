#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --skip-cpu --no-extras > gpurun_out/r2i_train.json 2> gpurun_out/r2i_train.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i_train.json').read()); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'probe', d['roofline']['ms_per_launch'])"; tail -2 gpurun_out/r2i_train.err
