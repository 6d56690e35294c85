"""ORACLE (test infrastructure only -- never imported by the product path).

Differentiable torch fp32 restatement of the GiLt loss, used as the checker of the ANALYTIC backward kernels
(torch autograd over this file + oracle/model_oracle.py plays the role of the reference's `loss.backward()`,
/root/reference/torchreid/engine/image/part_based_engine.py:98-103):

  * label-smoothed cross entropy      -> /root/reference/torchreid/losses/cross_entropy_loss.py:34-56
  * part-averaged batch-hard triplet  -> /root/reference/torchreid/losses/part_averaged_triplet_loss.py:35-195
  * GiLt weighted combination         -> /root/reference/torchreid/losses/GiLt_loss.py:45-119
  * body part attention loss          -> /root/reference/torchreid/losses/body_part_attention_loss.py:31-52 with
                                         /root/reference/torchreid/engine/image/part_based_engine.py:118-124

Parity pin: tests/test_oracle_vs_golden.py checks the VALUES against the numpy oracle and the reference's golden
losses, and the GRADIENTS against tests/golden/grads_hrnet32_k5.npz (minted from the unmodified reference modules by
oracle/make_golden.py grads).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS = "globl", "foreg", "conct", "parts"
DEFAULT_WEIGHTS = {GLOBAL: {"id": 1., "tr": 0.}, FOREGROUND: {"id": 1., "tr": 0.},
                   CONCAT_PARTS: {"id": 1., "tr": 0.}, PARTS: {"id": 0., "tr": 1.}}
FLT_MAX = torch.finfo(torch.float32).max


def cross_entropy_ls(logits, targets, eps=0.1):
    """cross_entropy_loss.py:34-49 (unweighted): (-t * log_softmax).mean(0).sum()."""
    n, c = logits.shape
    logp = F.log_softmax(logits, dim=1)
    t = torch.zeros_like(logp).scatter_(1, targets.view(-1, 1), 1.0)
    t = (1 - eps) * t + eps / c
    return (-t * logp).mean(0).sum()


def part_pairwise_dist(emb_knd, eps=1e-16):
    """part_averaged_triplet_loss.py:77-93."""
    dot = torch.matmul(emb_knd, emb_knd.transpose(2, 1))
    sq = dot.diagonal(dim1=1, dim2=2)
    d2 = F.relu(sq.unsqueeze(2) - 2 * dot + sq.unsqueeze(1))
    mask = (d2 == 0).float()
    return torch.sqrt(d2 + mask * eps) * (1.0 - mask)


def part_averaged_triplet(emb_nkd, labels, margin=0.3):
    """part_averaged_triplet_loss.py:35-65 (no visibility) + :95-195 (batch-hard mining, hard / soft margin)."""
    d = part_pairwise_dist(emb_nkd.transpose(1, 0)).mean(0)
    n = labels.shape[0]
    same = labels.view(1, -1) == labels.view(-1, 1)
    pos = same & ~torch.eye(n, dtype=torch.bool, device=d.device)
    neg = ~same
    hp = (d * pos.float() - (~pos).float()).max(1)[0]
    hn = (d * neg.float() + (~neg).float() * FLT_MAX).min(1)[0]
    valid = (hp != -1) & (hn != FLT_MAX)
    if valid.sum() == 0:
        return None
    hp, hn = hp[valid], hn[valid]
    if margin > 0:
        return F.relu(hp - hn + margin).mean()
    return torch.log1p(torch.exp(-(hn - hp))).mean()


def gilt(emb, ids, pids, weights=None, margin=0.3):
    """GiLt_loss.py:45-93 without visibility scores (mask_filtering_training False, the shipped default)."""
    w = DEFAULT_WEIGHTS if weights is None else weights
    total = 0.0
    for key in (GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS):
        if w[key]["id"] > 0:
            sc = ids[key]
            tg = pids
            if sc.dim() == 3:
                tg = pids.unsqueeze(1).expand(-1, sc.shape[1]).flatten(0, 1)
                sc = sc.flatten(0, 1)
            total = total + w[key]["id"] * cross_entropy_ls(sc, tg)
    for key in (GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS):
        if w[key]["tr"] > 0:
            e = emb[key] if emb[key].dim() == 3 else emb[key].unsqueeze(1)
            total = total + w[key]["tr"] * part_averaged_triplet(e, pids, margin)
    return total


def body_part_attention(pixel_logits, target_masks, label_smoothing=0.1):
    """part_based_engine.py:118-124 target preparation + body_part_attention_loss.py:41-47 ('cl')."""
    tm = F.interpolate(target_masks, pixel_logits.shape[2:], mode="bilinear", align_corners=True)
    tgt = tm.argmax(dim=1)
    kp1 = pixel_logits.shape[1]
    return F.cross_entropy(pixel_logits.permute(0, 2, 3, 1).flatten(0, 2), tgt.flatten(), label_smoothing=label_smoothing)
