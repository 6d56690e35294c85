#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tag in old new old new; do
  if [ $tag = old ]; then L="--lib tools/micro/lib_old.so"; else L=""; fi
  timeout 600 python bench.py --workload train --steps 10 --warmup 3 --skip-cpu --no-extras $L > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/ab_$tag.json').read()); print('$tag train', d['value'], d['ms_per_step'], 'probe', d['roofline']['ms_per_launch'])"
done
