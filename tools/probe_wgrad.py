"""Times the tcgen05 weight-gradient kernel (and the dgrad = forward kernel on flipped weights) on the HRNet layer shapes
at batch 256 (CUDA events, L2 flushed between launches)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpbreid_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
SHAPES = [(256, 64, 32, 32, 32, 3, 1)] if len(sys.argv) < 2 else [
    (256, 64, 32, 32, 32, 3, 1), (256, 32, 16, 64, 64, 3, 1), (256, 16, 8, 128, 128, 3, 1), (256, 8, 4, 256, 256, 3, 1),
    (256, 64, 32, 64, 64, 3, 1), (256, 64, 32, 64, 256, 1, 1), (256, 64, 32, 256, 64, 1, 1), (256, 128, 64, 64, 64, 3, 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for (n, h, w, cin, cout, k, st) in SHAPES:
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    up = 1 if st == 2 else 0
    dy = torch.randn(n, ho * (2 if up else 1), wo * (2 if up else 1), cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros(cout, k * k, cin, device=dev)
    call = lambda: _lib.call("bpb_conv_wgrad_bf16", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), n, h, w, cin, cout, k, st, up, s)
    for _ in range(2):
        call()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    byts = (n * h * w * cin + n * ho * wo * cout) * 2 + cout * k * k * cin * 4
    print(f"wgrad n{n} {h}x{w} {cin}->{cout} k{k} s{st}: {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s (algorithmic)")
