#!/bin/bash
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_n${N}_bench.json 2> gpurun_out/multi_n${N}_bench.err
echo rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/multi_n${N}_bench.json').read())
print('fwdloss N=${N}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('clocks'))
for k, v in d.get('extras', {}).items():
    print(' ', k, json.dumps(v)[:300])
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/multi_n${N}_bench.err | tail -5
