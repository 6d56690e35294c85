#!/bin/bash
# usage: gpu_r2_multi.sh N  -- distributed correctness check + default bench (with extras) + train bench on N GPUs
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/check_dist_train.py > gpurun_out/multi_n${N}_check.log 2>&1; tail -4 gpurun_out/multi_n${N}_check.log
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/multi_n${N}_bench.json 2> gpurun_out/multi_n${N}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/multi_n${N}_bench.json'))
print('fwdloss N=${N}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('clocks'))
for k, v in d.get('extras', {}).items():
    print(' ', k, json.dumps(v)[:300])
PY
tail -3 gpurun_out/multi_n${N}_bench.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload train --steps 10 --warmup 3 --skip-cpu --no-extras > gpurun_out/multi_n${N}_train.json 2> gpurun_out/multi_n${N}_train.err
python -c "
import json; d=json.load(open('gpurun_out/multi_n${N}_train.json')); print('train N=${N}', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"; tail -3 gpurun_out/multi_n${N}_train.err
