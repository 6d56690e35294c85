#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench.json').read())
print('fwdloss', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'conv probe ms', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])
for k, v in d.get('extras', {}).items():
    print(' ', k, json.dumps(v)[:420])
PY
tail -3 gpurun_out/r2_bench.err
