"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (bpbreid_b200/distributed.py) --
sharding that keeps identity groups together, rank-ordered all-gather, and "local ID terms + globally mined triplet
terms == the reference's whole-batch GiLt".  The arithmetic is delegated to the numpy oracle through a stub
criterion (the product losses are CUDA-only), so what is verified here is the exchange step itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bpbreid_b200 import distributed as D
from oracle import loss_oracle as LO
from tests import golden_inputs as GI


def test_shard_bounds_keep_identities_together():
    for gb, ws in [(256, 8), (256, 4), (64, 2), (48, 4)]:
        spans = [D.shard_bounds(gb, ws, r) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
        assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in spans)
    with pytest.raises(ValueError):
        D.shard_bounds(30, 2, 0)
    with pytest.raises(ValueError):                  # unequal shards would dead-lock all_gather_into_tensor
        D.shard_bounds(256, 3, 0)


class OracleCriterion:
    """GiLtLoss-shaped stub: same attributes / call signature, numbers from the numpy oracle."""

    def __init__(self, weights, use_visibility_scores=False):
        self.losses_weights = weights
        self.use_visibility_scores = use_visibility_scores

    def __call__(self, emb, vis, ids, pids):
        n = lambda d: {k: v.numpy() for k, v in d.items()}  # noqa: E731
        loss, summary = LO.gilt(n(emb), n(vis), n(ids), pids.numpy(), self.losses_weights, self.use_visibility_scores, 0.3)
        return torch.tensor(float(loss)), summary


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, case="market", use_vis=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = GI.loss_case(case)
        n = c["pids"].shape[0]
        lo, hi = D.shard_bounds(n, world, rank)
        sl = lambda d: {k: torch.from_numpy(v[lo:hi]) for k, v in d.items()}  # noqa: E731
        emb, vis, ids = sl(c["emb"]), sl(c["vis"]), sl(c["ids"])
        pids = torch.from_numpy(c["pids"][lo:hi])
        # gather order == rank order == original batch order
        g = D.all_gather_cat(emb["parts"])
        assert torch.equal(g, torch.from_numpy(c["emb"]["parts"]))
        gb = D.all_gather_cat(vis["parts"])
        assert gb.dtype == torch.from_numpy(c["vis"]["parts"]).dtype and torch.equal(gb, torch.from_numpy(c["vis"]["parts"]))
        loss, summary = D.distributed_gilt(OracleCriterion(c["weights"], use_vis), emb, vis, ids, pids)
        q.put((rank, float(loss), float(summary["parts"]["t"]), float(summary["globl"]["c"])))
    finally:
        dist.destroy_process_group()


def test_distributed_gilt_equals_whole_batch_gilt_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = GI.loss_case("market")
    whole, summ = LO.gilt(c["emb"], c["vis"], c["ids"], c["pids"], c["weights"], False, 0.3)
    for rank, loss, trip, _ in res:
        assert abs(loss - float(whole)) < 1e-5 * abs(float(whole)), (rank, loss, float(whole))
        assert abs(trip - summ["parts"]["t"]) < 1e-6          # mined over the gathered (global) batch
    # per-shard mining (what naive DDP would do) gives a different number: the exchange step is not optional
    lo, hi = D.shard_bounds(c["pids"].shape[0], world, 0)
    local = LO.part_averaged_triplet(c["emb"]["parts"][lo:hi], c["pids"][lo:hi], 0.3)
    assert abs(float(local[0]) - summ["parts"]["t"]) > 1e-4


@pytest.mark.parametrize("case", ["vis_bool", "vis_float"])
def test_distributed_gilt_with_visibility_scores_normalises_globally(case):
    """with visibility scores the identity terms are weighted means (bool: invisible rows dropped; float:
    sum(w*CE)/sum|w|, cross_entropy_loss.py:50-53).  Ranks see different weight totals, so the mean of the local means is
    NOT the whole-batch value: numerator and denominator must be reduced separately."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, case, True)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = GI.loss_case(case)
    whole, _ = LO.gilt(c["emb"], c["vis"], c["ids"], c["pids"], c["weights"], True, 0.3)
    for rank, loss, _, _ in res:
        assert abs(loss - float(whole)) < 1e-5 * abs(float(whole)), (rank, loss, float(whole))
    # the naive average of per-shard losses differs measurably
    lo, hi = D.shard_bounds(c["pids"].shape[0], world, 0)
    sl = lambda d, a, b: {k: v[a:b] for k, v in d.items()}  # noqa: E731
    only_id = {k: {"id": v["id"], "tr": 0.0} for k, v in c["weights"].items()}
    halves = [LO.gilt(sl(c["emb"], a, b), sl(c["vis"], a, b), sl(c["ids"], a, b), c["pids"][a:b], only_id, True, 0.3)[0]
              for a, b in ((lo, hi), (hi, c["pids"].shape[0]))]
    whole_id = LO.gilt(c["emb"], c["vis"], c["ids"], c["pids"], only_id, True, 0.3)[0]
    assert abs(float(np.mean(halves)) - float(whole_id)) > 1e-6
