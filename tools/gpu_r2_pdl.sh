#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
true
for tag in nopdl pdl nopdl pdl; do
  if [ $tag = nopdl ]; then L="--lib tools/micro/lib_nopdl.so"; else L=""; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --no-extras $L > gpurun_out/pdl_$tag.json 2> gpurun_out/pdl_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/pdl_$tag.json').read()); print('$tag fwdloss', d['value'], d['ms_per_step'], 'loss', d.get('loss'))"
  timeout 600 python bench.py --steps 20 --warmup 5 --batch 32 --skip-cpu --no-extras $L > gpurun_out/pdl32_$tag.json 2> gpurun_out/pdl32_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/pdl32_$tag.json').read()); print('$tag fwdloss bs32', d['value'], d['ms_per_step'])"
done
