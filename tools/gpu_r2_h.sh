#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/mask_cases_report.txt gpurun_out/fusion_report.txt
timeout 1500 python -m pytest tests/test_gpu_path.py tests/test_gpu_backward.py -m gpu -q -k "external_part or epilogue_fusion or bn_backward or every_conv" 2>&1 | tail -15 > gpurun_out/r2h_pytest.log; tail -8 gpurun_out/r2h_pytest.log; cat gpurun_out/fusion_report.txt
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --skip-cpu --no-extras > gpurun_out/r2f_train.json 2> gpurun_out/r2f_train.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_train.json')); print('train', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
