"""Times the two BatchNorm-backward passes on the HRNet map shapes at batch 256 (CUDA events; `cold` = L2 flushed before
every launch, `warm` = dY just written / re-read as inside the step)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpbreid_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:                       # A/B against another build of the library
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
dev = torch.device("cuda", 0)
SHAPES = [(256, 64, 32, 32), (256, 32, 16, 64), (256, 16, 8, 128), (256, 8, 4, 256), (256, 64, 32, 64), (256, 64, 32, 256)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
z = torch.zeros(64, device=dev)
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.call("bpb_memset_zero", z.data_ptr(), 256, s); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"timing floor (event pair around a 256-byte memset kernel): {float(np.median(ts))*1e3:.1f} us")
for (n, h, w, c) in SHAPES:
    raw = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    out = torch.empty_like(dy)
    mi = torch.stack([torch.zeros(c, device=dev), torch.ones(c, device=dev)]).contiguous()
    sc, sh = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    coef = torch.zeros(2, c, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    ws = torch.zeros(_lib.query("bpb_bn_bwd_workspace_bytes", c), dtype=torch.uint8, device=dev)
    red = lambda: _lib.call("bpb_bn_bwd_reduce", dy.data_ptr(), raw.data_ptr(), mi.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                            n, h, w, c, 0, 1, 0, ws.data_ptr(), coef.data_ptr(), dg.data_ptr(), db.data_ptr(), s)
    app = lambda: _lib.call("bpb_bn_bwd_apply", dy.data_ptr(), raw.data_ptr(), mi.data_ptr(), sc.data_ptr(), sh.data_ptr(),
                            coef.data_ptr(), n, h, w, c, 0, 1, 0, out.data_ptr(), s)
    res = []
    for call in (red, app):
        for cold in (True, False):
            for _ in range(2):
                call()
            ts = []
            for _ in range(7):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res.append(float(np.median(ts)) * 1e3)
    mb = n * h * w * c * 2 / 1e6
    # check against torch
    g = torch.where(raw.float() > 0, dy.float(), torch.zeros((), device=dev))
    ref_db = g.sum((0, 1, 2)); ref_dg = (g * raw.float()).sum((0, 1, 2))
    e1 = float((db - ref_db).abs().max() / ref_db.abs().max()); e2 = float((dg - ref_dg).abs().max() / ref_dg.abs().max())
    print(f"bn bwd n{n} {h}x{w} C={c} ({mb:.1f} MB map): reduce cold {res[0]:6.1f} us warm {res[1]:6.1f} us ({2*mb/res[0]:.2f} TB/s cold) | "
          f"apply cold {res[2]:6.1f} us warm {res[3]:6.1f} us ({3*mb/res[2]:.2f} TB/s cold) | err {e1:.1e} {e2:.1e}")
