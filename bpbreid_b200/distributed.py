"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed; NCCL on B200, gloo in CPU tests).

The batch shards along the image axis; the model is replicated and BatchNorm statistics stay per replica
(= the reference's nn.DataParallel behaviour, torchreid/scripts/main.py:256-257; no SyncBN anywhere).  The reference
gathers the outputs onto GPU 0 and computes GiLt over the WHOLE batch (torchreid/engine/image/part_based_engine.py:82-95).
Here that becomes two exchange steps in forward + loss:

  (1) one all-gather of the embeddings that feed the triplet terms (+ person ids, + visibility when used); every rank
      then computes only ITS anchor rows of the N_global x N_global part-distance matrix (bpb_triplet_fwd_rows) and
      mines over all columns, so the mined pairs are exactly the reference's and the N^2 work is split across ranks;
  (2) one all-reduce of a small vector of partial sums: the triplet {sum loss, sum trivial, n_valid, rows} and, for the
      identity terms, numerator sum(w_i * CE_i) and denominator sum(|w_i|) (or the row count), so that the global loss
      is normalised ONCE like the reference (cross_entropy_loss.py:50-53), not as a mean of per-rank means.

Training adds (3) an all-reduce of the triplet gradient w.r.t. the gathered embeddings and (4) the all-reduce of the
flat parameter-gradient buffer (bpbreid_b200.train).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def is_dist(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def shard_bounds(global_batch: int, world_size: int, rank: int, num_instances: int = 4):
    """Contiguous shard [lo, hi) of the global batch that keeps the P x K sampler's identity groups together
    (RandomIdentitySampler emits `num_instances` consecutive samples per identity, torchreid/data/sampler.py:11-76).
    Shards must be equal (the collectives and the static per-rank plans need one shape): a global batch whose identity
    groups do not divide evenly over the ranks is rejected."""
    if global_batch % num_instances:
        raise ValueError("global batch must be a multiple of num_instances")
    groups = global_batch // num_instances
    if groups % world_size:
        raise ValueError(f"{groups} identity groups do not split evenly over {world_size} ranks: pad the batch or change "
                         "the batch size (unequal shards would dead-lock all_gather_into_tensor)")
    per = groups // world_size
    return rank * per * num_instances, (rank + 1) * per * num_instances


def all_gather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equal-shaped shards along dim 0 in rank order."""
    if not is_dist(group):
        return t
    ws = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if t.dtype is torch.bool:
        dist.all_gather_into_tensor(out.view(torch.uint8), t.view(torch.uint8), group=group)
    else:
        dist.all_gather_into_tensor(out, t, group=group)
    return out


def gather_for_mining(embeddings: dict, visibility: dict, pids: torch.Tensor, keys, group=None):
    """All-gather the tensors the triplet terms need (`keys`, e.g. ['parts'] or ['parts', 'foreg'])."""
    g_emb = {k: all_gather_cat(embeddings[k], group) for k in keys}
    g_vis = {k: all_gather_cat(visibility[k], group) for k in keys}
    return g_emb, g_vis, all_gather_cat(pids, group)


def sum_across_ranks(x: torch.Tensor, group=None) -> torch.Tensor:
    if not is_dist(group):
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y


def mean_across_ranks(x: torch.Tensor, group=None) -> torch.Tensor:
    if not is_dist(group):
        return x
    return sum_across_ranks(x, group) / dist.get_world_size(group)


class ShardedTriplet:
    """Part-averaged batch-hard triplet over the GLOBAL batch with per-rank anchor rows (static buffers, CUDA only).

    local embeddings [K, n, D] fp32 (part-major, as the forward plan stores them) -> all-gather -> [W, K, n, D];
    forward(): mined loss over the global batch, identical on every rank; backward(): gradient w.r.t. the LOCAL
    embeddings (all-reduced over the ranks that mined pairs touching them)."""

    def __init__(self, n, K, D, margin, device, group=None):
        self.n, self.K, self.D, self.margin, self.group = n, K, D, float(margin), group
        self.world = dist.get_world_size(group) if is_dist(group) else 1
        self.rank = dist.get_rank(group) if is_dist(group) else 0
        W, N = self.world, self.world * n
        self.N = N
        self.emb_g = torch.empty(W, K, n, D, device=device)
        self.pids_g = torch.empty(N, dtype=torch.int64, device=device)
        self.dist_rows = torch.empty(n, N, device=device)
        self.hp, self.hn = torch.empty(n, device=device), torch.empty(n, device=device)
        self.hp_idx = torch.empty(n, device=device, dtype=torch.int32)
        self.hn_idx = torch.empty(n, device=device, dtype=torch.int32)
        self.sums = torch.zeros(4, device=device)
        self.out4 = torch.zeros(4, device=device)
        self.demb_g = torch.zeros(W, K, n, D, device=device)

    # The step is split into kernel phases and NCCL phases so that a caller (bpbreid_b200.train.TrainStep) can capture the
    # kernel phases in CUDA graphs and issue the collectives between them; forward() / backward() run them back to back.
    def gather(self, emb_knd_local: torch.Tensor, pids_local: torch.Tensor):
        if self.world > 1:
            dist.all_gather_into_tensor(self.emb_g.view(-1), emb_knd_local.reshape(-1), group=self.group)
            dist.all_gather_into_tensor(self.pids_g, pids_local.contiguous(), group=self.group)
        else:
            self.emb_g.view(-1).copy_(emb_knd_local.reshape(-1))
            self.pids_g.copy_(pids_local)

    def mine(self):
        _lib.call("bpb_triplet_fwd_rows", self.emb_g.data_ptr(), self.pids_g.data_ptr(), None, 0, self.N, self.K, self.D,
                  self.margin, self.rank * self.n, self.n, self.n, 1, self.dist_rows.data_ptr(), self.hp.data_ptr(),
                  self.hn.data_ptr(), self.hp_idx.data_ptr(), self.hn_idx.data_ptr(), self.sums.data_ptr(),
                  _lib.stream_ptr())

    def reduce_stats(self):
        if self.world > 1:
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        _lib.call("bpb_triplet_finish", self.sums.data_ptr(), self.out4.data_ptr(), _lib.stream_ptr())

    def forward(self, emb_knd_local: torch.Tensor, pids_local: torch.Tensor):
        self.gather(emb_knd_local, pids_local)
        self.mine()
        self.reduce_stats()
        self.finish()
        return self.out4

    def bwd_kernel(self, loss_weight: float):
        self.demb_g.zero_()
        _lib.call("bpb_triplet_bwd_rows", self.emb_g.data_ptr(), None, 0, self.hp.data_ptr(), self.hn.data_ptr(),
                  self.hp_idx.data_ptr(), self.hn_idx.data_ptr(), self.out4.data_ptr(), self.N, self.K, self.D, self.margin,
                  float(loss_weight), self.rank * self.n, self.n, self.n, self.demb_g.data_ptr(), _lib.stream_ptr())

    def bwd_reduce(self):
        if self.world > 1:
            dist.all_reduce(self.demb_g, op=dist.ReduceOp.SUM, group=self.group)

    def bwd_copy(self, d_emb_local: torch.Tensor):
        d_emb_local.view(self.K, self.n, self.D).copy_(self.demb_g[self.rank])

    def backward(self, loss_weight: float, d_emb_local: torch.Tensor):
        """d_emb_local [K, n, D] = gradient of loss_weight * loss (the global loss; every rank holds the same value).
        OVERWRITES d_emb_local (a device-to-device copy of this rank's shard of the all-reduced gradient): call it before
        the other contributions are accumulated."""
        self.bwd_kernel(loss_weight)
        self.bwd_reduce()
        self.bwd_copy(d_emb_local)
        return d_emb_local


def distributed_gilt(criterion, embeddings, visibility, id_scores, pids, group=None, cache=None):
    """GiLt over the global batch (forward values).  With world_size == 1 this is exactly criterion(...).

    Identity terms: every rank evaluates its shard; numerator and denominator of the (weighted) mean are all-reduced
    separately and divided once.  Triplet terms: ShardedTriplet (CUDA) -- or, on CPU tensors (gloo tests with a stub
    criterion), the gathered batch through the criterion itself."""
    from .constants import GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS
    if not is_dist(group):
        return criterion(embeddings, visibility, id_scores, pids)
    world = dist.get_world_size(group)
    w = criterion.losses_weights
    keys = (GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS)
    tr_keys = [k for k in keys if w[k]["tr"] > 0]
    id_keys = [k for k in keys if w[k]["id"] > 0]
    use_vis = bool(getattr(criterion, "use_visibility_scores", False))
    saved = criterion.losses_weights
    total = None
    summary = {}
    try:
        # ---- identity terms, one key at a time so that each can be re-normalised globally
        for k in id_keys:
            criterion.losses_weights = {kk: {"id": (1.0 if kk == k else 0.0), "tr": 0.0} for kk in saved}
            local, s_k = criterion(embeddings, visibility, id_scores, pids)
            summary.setdefault(k, {}).update(s_k.get(k, {}))
            vis = visibility[k]
            if use_vis:
                # local value = sum(w*CE)/max(sum|w|,eps): recover the numerator, reduce numerator and denominator
                den = vis.float().abs().sum().reshape(1).clamp_min(1e-12)
            else:
                rows = id_scores[k].shape[0] * (id_scores[k].shape[1] if id_scores[k].dim() == 3 else 1)
                den = torch.full((1,), float(rows), device=local.device)
            pair = torch.stack([local.reshape(()) * den[0], den[0]])
            pair = sum_across_ranks(pair, group)
            term = w[k]["id"] * pair[0] / pair[1].clamp_min(1e-12)
            total = term if total is None else total + term
        # ---- triplet terms over the global batch
        for k in tr_keys:
            e = embeddings[k] if embeddings[k].dim() == 3 else embeddings[k].unsqueeze(1)
            if e.is_cuda and not use_vis:
                n, K, D = e.shape
                key = (k, n, K, D)
                st = None if cache is None else cache.get(key)
                if st is None:
                    st = ShardedTriplet(n, K, D, criterion.part_triplet_loss.margin, e.device, group)
                    if cache is not None:
                        cache[key] = st
                out4 = st.forward(e.permute(1, 0, 2).contiguous().float(), pids.to(torch.int64))
                summary.setdefault(k, {}).update({"t": out4[0], "tt": out4[1], "vt": out4[2]})
                term = w[k]["tr"] * out4[0]
            else:
                g_emb, g_vis, g_pids = gather_for_mining(embeddings, visibility, pids, [k], group)
                criterion.losses_weights = {kk: {"id": 0.0, "tr": (saved[kk]["tr"] if kk == k else 0.0)} for kk in saved}
                term, s_k = criterion(dict(embeddings, **g_emb), dict(visibility, **g_vis), id_scores, g_pids)
                summary.setdefault(k, {}).update(s_k.get(k, {}))
            total = term if total is None else total + term
    finally:
        criterion.losses_weights = saved
    return total, summary
