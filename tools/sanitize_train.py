"""One fused training step under compute-sanitizer (memcheck): python tools/sanitize_train.py N"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bpbreid_b200 import losses, train  # noqa: E402
from tests import golden_inputs as GI  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c = GI.model_case("hrnet32_k5_ctr", n=n)
m = c["model"]
m.load_state_dict(c["state_dict"])
m = m.to("cuda")
m.materialize_spatial_features = False
ts = train.TrainStep(m, losses.GiLtLoss(None, use_gpu=True), None, bpa_weight=0.35, use_graph=False)
dev = "cuda"
plan = ts.forward_backward(torch.from_numpy(c["images"]).to(dev), torch.from_numpy(c["pids"]).to(dev),
                           torch.from_numpy(c["masks"]).to(dev))
torch.cuda.synchronize()
print("loss", float(plan.loss_out[0]))
ts._optimizer(plan)
torch.cuda.synchronize()
print("step ok")
