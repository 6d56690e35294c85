"""2 ranks: the data-parallel training step (sharded global-batch mining + bucketed gradient all-reduce) must produce
the same gradients as ONE rank stepping on the concatenated batch with the SAME per-shard BatchNorm statistics...
which a single process cannot reproduce exactly (BN statistics are per replica), so the checks are:
  (a) both ranks end with identical gradients (all-reduce consistency) and identical triplet loss,
  (b) the gathered mining loss equals the single-process loss on the gathered embeddings,
  (c) parameters stay identical across ranks after an optimizer step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpbreid_b200 import losses, train  # noqa: E402
from tests import golden_inputs as GI  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c = GI.model_case("hrnet32_k5_ctr", n=16)
m = c["model"]
m.load_state_dict(c["state_dict"])
m = m.to(dev)
m.materialize_spatial_features = False
crit = losses.GiLtLoss(None, use_gpu=True)
ts = train.TrainStep(m, crit, None, use_graph=True)
n = 16 // world
lo, hi = rank * n, (rank + 1) * n
imgs = torch.from_numpy(c["images"][lo:hi]).to(dev)
pids = torch.from_numpy(c["pids"][lo:hi]).to(dev)
plan = ts.forward_backward(imgs, pids)
torch.cuda.synchronize()
g = ts.flat.g.clone()
gs = [torch.empty_like(g) for _ in range(world)]
dist.all_gather(gs, g)
same = all(torch.equal(gs[0], x) for x in gs[1:])
st = plan.sharded[0][0]
# (b) single-process triplet on the gathered embeddings
emb_all = st.emb_g.permute(1, 0, 2, 3).reshape(st.K, st.N, st.D).permute(1, 0, 2).contiguous()
tl = losses.PartAveragedTripletLoss(margin=0.3)
ref = tl(emb_all, st.pids_g)
torch.cuda.synchronize()
tri_first, valid_first = float(st.out4[0]), float(st.out4[3])
for _ in range(3):                                        # phase graphs captured on the 2nd call, replayed afterwards
    loss = ts.step(imgs, pids)
torch.cuda.synchronize()
p = ts.flat.p.clone()
ps = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(ps, p)
same_p = all(torch.equal(ps[0], x) for x in ps[1:])
if rank == 0:
    print(f"grads identical across ranks: {same}; |g| = {float(g.norm()):.4f}")
    print(f"sharded triplet loss {tri_first:.6f} vs single-process on gathered embeddings {float(ref[0]):.6f}; "
          f"valid {valid_first}")
    print(f"params identical across ranks after Adam: {same_p}; loss {float(loss):.5f}")
    assert same and same_p and abs(tri_first - float(ref[0])) < 1e-5 * abs(float(ref[0]))
    print("dist train ok")
dist.destroy_process_group()
