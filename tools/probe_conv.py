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
                             sc.data_ptr(), sh.data_ptr(), s)
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
    if os.environ.get("BPB_CONV_ROLES"):
        # per-role cycle counters of CTA 0 (diagnostic hook)
        prof = torch.zeros(128 + 2400, dtype=torch.int64, device=dev)
        _lib.call("bpb_conv_set_profile_buffer", prof.data_ptr())
        call(); torch.cuda.synchronize()
        _lib.call("bpb_conv_set_profile_buffer", None)
        full = prof.cpu().numpy()
        pr = full[:128].reshape(16, 8)
        tl = full[128:].reshape(600, 4)
        tl = tl[tl[:, 0] > 0]
        if len(tl):
            t0 = tl[:, 0].min()
            st, en = tl[:, 0] - t0, tl[:, 1] - t0
            per_sm = np.bincount(tl[:, 2].astype(int))
            print(f"    kernel span {en.max()/1e3:.1f} us  CTAs {len(tl)}  SMs used {np.count_nonzero(per_sm)}  max CTAs/SM {per_sm.max()}  start ns [min {st.min()} med {int(np.median(st))} max {st.max()}]  end ns [min {en.min()} med {int(np.median(en))} max {en.max()}]")
        print(f"    cta0 cycles: setup {pr[15,0]}  producer-done {pr[15,1]}  all-roles-done {pr[15,2]}")
        print(f"    producer: total {pr[0,0]}  wait_empty {pr[0,1]}  issue {pr[0,2]}")
        for wv in range(3):
            if pr[1 + wv, 0]:
                print(f"    mma{wv}: total {pr[1+wv,0]}  wait_acc_empty {pr[1+wv,1]}  wait_full {pr[1+wv,2]}  issue {pr[1+wv,3]}  wait_b_resident {pr[1+wv,4]}")
        for gq in range(2):
            if pr[5 + gq, 0]:
                print(f"    epilogue{gq}: total {pr[5+gq,0]}  wait_acc_full {pr[5+gq,1]}  drain {pr[5+gq,2]}  stats {pr[5+gq,3]}  store {pr[5+gq,4]}  lt {pr[5+gq,5]}")
