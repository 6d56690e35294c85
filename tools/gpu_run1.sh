#!/bin/bash
# first GPU bring-up: isolate risky tests in separate processes so a sticky CUDA error does not hide the rest
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
L=gpurun_out/run1.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python -m pytest -q --tb=short -p no:cacheprovider "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
run tests/test_gpu_kernels.py -k "test_linear_tf32"
run tests/test_gpu_kernels.py -k "conv_tc and 64, 32, 32, 32, 3, 1"
run tests/test_gpu_kernels.py -k "conv_tc and 32, 16, 64, 64, 3, 1"
run tests/test_gpu_kernels.py -k "conv_tc and 3, 2"
run tests/test_gpu_kernels.py -k "conv_tc and not (3, 2)"
run tests/test_gpu_kernels.py -k "not conv_tc and not linear"
run tests/test_gpu_path.py -k "gilt or triplet or body_part"
run tests/test_gpu_path.py -k "simt"
run tests/test_gpu_path.py -k "tcgen05"
run tests/test_gpu_path.py -k "rank or signature"
run tests/test_gpu_path.py -k "medium_scale or full_size"
run tests/test_gpu_path.py -k "model_forward or conv_paths"
grep -E "^(===|rc=|FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed" $L | tail -80
echo "=== bench" >> $L
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench1.log
