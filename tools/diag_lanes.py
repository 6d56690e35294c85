"""Diagnostic: does the stream-lane schedule reproduce the single-stream schedule under CUDA-graph replay?
Runs eval-mode HRNet (seeded weights) in four configurations and reports the first diverging block (debug trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import golden_inputs as GI

DEV = torch.device("cuda", 0)


def run(single, graph, debug, reps=3):
    os.environ["BPB_SINGLE_STREAM"] = "1" if single else "0"
    c = GI.model_case("hrnet32_k5")
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to(DEV)
    m.train(False)
    m.use_cuda_graph = graph
    m.debug_trace = debug
    x = torch.from_numpy(c["images"]).to(DEV)
    outs = []
    with torch.no_grad():
        for _ in range(reps):
            emb, vis, ids, pix, _, masks = m(x)
            torch.cuda.synchronize()
            outs.append([t.float().clone() for t in (emb["parts"], emb["bn_globl"], ids["conct"], pix)])
    pl = next(iter(m._plans.values()))
    trace = [(t["kind"], t["names"], [o.t.float().clone() for o in t["outs"]]) for t in pl.trace] if debug else None
    return outs, trace


def rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


ref, ref_tr = run(True, False, True, 1)
for single, graph, debug in ((False, False, False), (False, True, False), (True, True, False), (False, True, True)):
    outs, tr = run(single, graph, debug)
    print(f"single={single} graph={graph} debug={debug}:",
          " ".join(f"rep{i}:" + ",".join(f"{rel(a, b):.1e}" for a, b in zip(o, ref[0])) for i, o in enumerate(outs)))
    if tr is not None:
        for (k, names, os_), (k2, names2, os2) in zip(tr, ref_tr):
            d = max(rel(a, b) for a, b in zip(os_, os2))
            if d > 1e-5:
                print("   first diverging block:", k, names, f"{d:.2e}")
                break
        else:
            print("   debug trace identical to the single-stream eager trace")
