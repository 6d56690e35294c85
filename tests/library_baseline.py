"""GPU *library* baseline (not collected by pytest; run as a script on a B200): the oracle's functional PyTorch restatement
of the reference forward, executed by cuDNN / cuBLAS through ATen -- the reference ships no kernels of its own, so this is
"what a user of the reference gets" on this GPU (last row).  Forward only, no_grad, batch 256,
256x128, K=5, HRNet-W32, train-mode BN, (a) fp32 NCHW, (b) bf16 autocast, (c) bf16 autocast + channels_last.
The oracle is test infrastructure: this script only times it, nothing in the product imports it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import model_oracle  # noqa: E402
from tests.golden_inputs import build_product_model  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n = 256
    m = build_product_model("hrnet32", 5)
    sd = {k: v.to(dev) for k, v in m.state_dict().items()}
    torch.manual_seed(0)
    x = torch.randn(n, 3, 256, 128, device=dev)
    out = {}

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    with torch.no_grad():
        ms = timeit(lambda: model_oracle.bpbreid_forward(sd, x, "hrnet32", 5, training=True))
        out["fp32_nchw"] = {"ms": ms, "images_per_s": n / ms * 1e3}
        def bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return model_oracle.bpbreid_forward(sd, x, "hrnet32", 5, training=True)
        ms = timeit(bf16)
        out["bf16_autocast"] = {"ms": ms, "images_per_s": n / ms * 1e3}
        xc = x.contiguous(memory_format=torch.channels_last)
        sdc = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
        def bf16_cl():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return model_oracle.bpbreid_forward(sdc, xc, "hrnet32", 5, training=True)
        ms = timeit(bf16_cl)
        out["bf16_autocast_channels_last"] = {"ms": ms, "images_per_s": n / ms * 1e3}
    out["note"] = "forward only (no loss), ATen eager; torch %s, cudnn %s" % (torch.__version__, torch.backends.cudnn.version())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
