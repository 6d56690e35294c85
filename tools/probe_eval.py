"""Times the evaluation kernels at BASELINE.json config 5 (4096 q x 16384 g, P=6, D=512): part distance (tcgen05 /
SIMT), ranking without sort, stable argsort; CUDA events, L2 flushed between launches.  Prints algorithmic GB/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bpbreid_b200 import _lib, metrics

dev = torch.device("cuda", 0)
nq, ng, p, d = 4096, 16384, 6, 512
g = torch.Generator(device=dev).manual_seed(0)
qf = torch.nn.functional.normalize(torch.randn(nq, p, d, device=dev, generator=g), dim=-1)
gf = torch.nn.functional.normalize(torch.randn(ng, p, d, device=dev, generator=g), dim=-1)
qv = torch.rand(nq, p, device=dev, generator=g) > 0.15
gv = torch.rand(ng, p, device=dev, generator=g) > 0.15
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dist = torch.empty(nq, ng, device=dev)
part = torch.empty(p, nq, ng, device=dev)
scratch = torch.empty(_lib.query("bpb_part_distance_scratch_floats", nq, ng, p), device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = {}
in_bytes = (nq + ng) * p * d * 4
full_bytes = in_bytes + nq * ng * 4 * (p + 1)
comb_bytes = in_bytes + nq * ng * 4
for name, impl, want in (("tc_full", 0, True), ("tc_combined_only", 0, False), ("simt_full", 1, True)):
    ms = timeit(lambda: metrics.part_distance_device(qf, gf, qv, gv, "mean", "euclidean", impl, want,
                                                     out=(dist, part if want else None, scratch)), 3 if impl else 5)
    b = full_bytes if want else comb_bytes
    out[name] = dict(ms=ms, algorithmic_GB=b / 1e9, GBps=b / ms / 1e6, TFLOPs=2.0 * p * nq * ng * d / ms / 1e9)
    print(name, out[name])
qp = torch.randint(0, 1500, (nq,), device=dev); gp = torch.randint(0, 1500, (ng,), device=dev)
qc = torch.randint(0, 6, (nq,), device=dev); gc = torch.randint(0, 6, (ng,), device=dev)
ms = timeit(lambda: metrics.rank_device(dist, qp, gp, qc, gc, 50))
out["rank_no_sort"] = dict(ms=ms, GBps=nq * ng * 4 / ms / 1e6)
print("rank_no_sort", out["rank_no_sort"])
sub = dist[:1024].contiguous()
ms = timeit(lambda: metrics.argsort_rows(sub))
out["argsort_1024rows"] = dict(ms=ms)
print("argsort_1024rows", out["argsort_1024rows"])
json.dump(out, open("gpurun_out/probe_eval.json", "w"), indent=1)
