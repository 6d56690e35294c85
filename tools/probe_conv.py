"""Times the tcgen05 conv kernel on the HRNet layer shapes at batch 256 (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bpbreid_b200 import _lib

dev = torch.device("cuda", 0)
SHAPES = [(256, 64, 32, 32, 32, 3, 1)] if len(sys.argv) < 2 else [
    (256, 64, 32, 32, 32, 3, 1), (256, 32, 16, 64, 64, 3, 1), (256, 16, 8, 128, 128, 3, 1), (256, 8, 4, 256, 256, 3, 1),
    (256, 64, 32, 64, 64, 3, 1), (256, 64, 32, 64, 256, 1, 1), (256, 64, 32, 256, 64, 1, 1), (256, 64, 32, 256, 32, 3, 1),
    (256, 64, 32, 32, 128, 1, 1), (256, 8, 4, 256, 1024, 1, 1), (256, 128, 64, 64, 64, 3, 2), (256, 64, 32, 32, 32, 3, 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for (n, h, w, cin, cout, k, st) in SHAPES:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wt = (torch.randn(cout, k * k * cin, device=dev) * 0.05).to(torch.bfloat16)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // st + 1, (w + 2 * pad - k) // st + 1
    y = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.bfloat16)
    ws = torch.zeros(_lib.query("bpb_conv_bn_workspace_bytes", cout), dtype=torch.uint8, device=dev)
    g, b = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    rm, rv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
    sc, sh = torch.empty(cout, device=dev), torch.empty(cout, device=dev)
    call = lambda: _lib.call("bpb_conv_bn_bf16", x.data_ptr(), wt.data_ptr(), y.data_ptr(), None, n, h, w, cin, cout, k, st, 0,
                             ws.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(), 0,
                             sc.data_ptr(), sh.data_ptr(), None, s)
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
    byts = (n * h * w * cin + n * ho * wo * cout + cout * k * k * cin) * 2
    print(f"conv n{n} {h}x{w} {cin}->{cout} k{k} s{st}: {ms*1e3:8.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s (algorithmic)")
