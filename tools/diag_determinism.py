"""Diagnostic: where does run-to-run variation of the training step start?  Serial recording, debug mode (every forward
activation and backward record kept): run the step twice and report the first forward block / backward record whose
buffers are not bit-identical between the two runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BPB_SINGLE_STREAM"] = "1"
import torch
from bpbreid_b200 import losses, train
from tests import golden_inputs as GI

DEV = torch.device("cuda", 0)
c = GI.model_case("hrnet32_k5_ctr", n=8)
m = c["model"]
m.load_state_dict(c["state_dict"])
m = m.to(DEV)
m.materialize_spatial_features = False
crit = losses.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, use_gpu=True)
ts = train.TrainStep(m, crit, None, bpa_weight=0.35, use_graph=False)
ts.debug = True
args = [torch.from_numpy(c[k]).to(DEV) for k in ("images", "pids", "masks")]


def snap(plan):
    fwd = [(t["kind"], t["names"], [o.t.clone() for o in t["outs"]]) for t in plan.trace]
    bwd = [(r["name"], r["dy"].t.clone(), r["draw"].t.clone(), None if r["dx"] is None else r["dx"].t.clone())
           for r in plan.btrace]
    hd = [d.t.clone() for d in plan.head_dx]
    def cl(v):
        if isinstance(v, (list, tuple)):
            return [cl(x) for x in v]
        return (v.t if hasattr(v, "t") and not torch.is_tensor(v) else v).clone()
    dbg = {k: cl(v) for k, v in plan.head_dbg.items()}
    return fwd, bwd, hd, ts.flat.g.clone(), plan.loss_out.clone(), dbg


plan = ts.forward_backward(*args); torch.cuda.synchronize()
a = snap(plan)
plan = ts.forward_backward(*args); torch.cuda.synchronize()
b = snap(plan)
print("loss bit-identical:", torch.equal(a[4], b[4]), a[4].tolist(), b[4].tolist())
nf = 0
for (k, names, o1), (_, _, o2) in zip(a[0], b[0]):
    if not all(torch.equal(x, y) for x, y in zip(o1, o2)):
        d = max(float((x.float() - y.float()).abs().max()) for x, y in zip(o1, o2))
        if nf < 5:
            print("forward block differs:", k, names, f"max abs {d:.3e}")
        nf += 1
print(f"forward blocks differing: {nf} of {len(a[0])}")
def rn(x, y):
    return float((x.float() - y.float()).norm() / (y.float().norm() + 1e-30))


print("head dx rel diff:", [f"{rn(x, y):.2e}" for x, y in zip(a[2], b[2])])
hd1, hd2 = a[5], b[5]
for k in hd1:
    print("  head dbg", k, [f"{rn(x, y):.2e}" for x, y in zip(hd1[k], hd2[k])] if isinstance(hd1[k], (list, tuple))
          else f"{rn(hd1[k], hd2[k]):.2e}")
print("dy rel diff along the backward:", " ".join(f"{i}:{rn(a[1][i][1], b[1][i][1]):.1e}" for i in range(0, len(a[1]), 16)))
nb = 0
for i, ((name, dy1, dr1, dx1), (_, dy2, dr2, dx2)) in enumerate(zip(a[1], b[1])):
    e = [torch.equal(dy1, dy2), torch.equal(dr1, dr2), dx1 is None or torch.equal(dx1, dx2)]
    if not all(e):
        if nb < 8:
            print(f"backward record {i} ({name}) differs: dy {e[0]} draw {e[1]} dx {e[2]}")
        nb += 1
print(f"backward records differing: {nb} of {len(a[1])}")
g1, g2 = a[3], b[3]
print("flat grad rel diff", float((g1 - g2).norm() / g2.norm()))
