#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "wgrad or bn_backward or every_conv" 2>&1 | tail -12
python tools/probe_wgrad.py all 2>&1 | tail -9
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --skip-cpu --no-extras > gpurun_out/r2f_train.json 2> gpurun_out/r2f_train.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_train.json')); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'conv probe us', d['roofline']['ms_per_launch']*1e3)"
tail -3 gpurun_out/r2f_train.err
