#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_n1_same_box.log 2>&1; echo "bench n1 rc=$?"; tail -1 gpurun_out/bench_n1_same_box.log | cut -c1-200
