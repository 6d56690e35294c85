"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed; NCCL on B200, gloo in CPU tests).

The batch shards along the image axis; the model is replicated and BatchNorm statistics stay per replica
(= the reference's nn.DataParallel behaviour, torchreid/scripts/main.py:256-257; no SyncBN anywhere).  The only
exchange step of forward+loss is the one the reference performs implicitly when DataParallel gathers the outputs
onto GPU 0 and computes GiLt over the *whole* batch (torchreid/engine/image/part_based_engine.py:82-95): an
all-gather of the embeddings that feed the triplet terms, their visibility scores and the person ids, so that
batch-hard mining sees every sample of the global batch.  The identity (cross-entropy) terms are per-sample means:
they are computed on the local shard and averaged across ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world_size: int, rank: int, num_instances: int = 4):
    """Contiguous shard [lo, hi) of the global batch that keeps the P x K sampler's identity groups together
    (RandomIdentitySampler emits `num_instances` consecutive samples per identity, torchreid/data/sampler.py:11-76)."""
    if global_batch % num_instances:
        raise ValueError("global batch must be a multiple of num_instances")
    groups = global_batch // num_instances
    base, rem = divmod(groups, world_size)
    lo_g = rank * base + min(rank, rem)
    hi_g = lo_g + base + (1 if rank < rem else 0)
    return lo_g * num_instances, hi_g * num_instances


def all_gather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate equal-shaped shards along dim 0 in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
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


def mean_across_ranks(x: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
    return y / dist.get_world_size(group)


def distributed_gilt(criterion, embeddings, visibility, id_scores, pids, group=None):
    """GiLt over the global batch: local identity terms (averaged over ranks) + triplet terms mined over the
    all-gathered embeddings.  With world_size == 1 this is exactly criterion(...)."""
    from .constants import GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return criterion(embeddings, visibility, id_scores, pids)
    w = criterion.losses_weights
    tr_keys = [k for k in (GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS) if w[k]["tr"] > 0]
    id_only = {k: {"id": w[k]["id"], "tr": 0.} for k in w}
    tr_only = {k: {"id": 0., "tr": w[k]["tr"]} for k in w}
    saved = criterion.losses_weights
    try:
        criterion.losses_weights = id_only
        id_loss, summary = criterion(embeddings, visibility, id_scores, pids)
        id_loss = mean_across_ranks(id_loss.reshape(1), group)[0]
        g_emb, g_vis, g_pids = gather_for_mining(embeddings, visibility, pids, tr_keys, group)
        criterion.losses_weights = tr_only
        full_emb = dict(embeddings, **g_emb)
        full_vis = dict(visibility, **g_vis)
        tr_loss, tr_summary = criterion(full_emb, full_vis, id_scores, g_pids)
    finally:
        criterion.losses_weights = saved
    for k, v in tr_summary.items():
        summary.setdefault(k, {}).update(v)
    return id_loss + tr_loss, summary
