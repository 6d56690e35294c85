"""ORACLE (test infrastructure only -- never imported by the product path).

numpy fp32 restatement of the loss side of the hot path:

  * label-smoothed cross entropy      -> /root/reference/torchreid/losses/cross_entropy_loss.py:34-56
  * part-averaged batch-hard triplet  -> /root/reference/torchreid/losses/part_averaged_triplet_loss.py:35-224
    (masked mean with the -1 sentinel -> /root/reference/torchreid/utils/tensortools.py:10-19)
  * GiLt weighted combination         -> /root/reference/torchreid/losses/GiLt_loss.py:45-119
  * body part attention loss          -> /root/reference/torchreid/losses/body_part_attention_loss.py:31-52
    (+ target preparation             -> /root/reference/torchreid/engine/image/part_based_engine.py:118-124)

Everything is computed in float32 with the same operation order as the reference's tensor expressions so that
exact-zero / sentinel comparisons (d == 0, d == -1, d == FLT_MAX) behave identically.
Parity pin: tests/test_oracle_vs_golden.py (fixtures minted from the reference by oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
FLT_MAX = np.finfo(np.float32).max

GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS, PIXELS = "globl", "foreg", "conct", "parts", "pixls"
DEFAULT_WEIGHTS = {GLOBAL: {"id": 1., "tr": 0.}, FOREGROUND: {"id": 1., "tr": 0.},
                   CONCAT_PARTS: {"id": 1., "tr": 0.}, PARTS: {"id": 0., "tr": 1.}}


def log_softmax(x):
    x = np.asarray(x, F32)
    m = x.max(axis=1, keepdims=True)
    z = x - m
    return z - np.log(np.exp(z).sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


def cross_entropy_ls(logits, targets, eps=0.1, weights=None):
    """cross_entropy_loss.py:34-56: t = (1-eps)*onehot + eps/C ; (-t*logp).mean(0).sum(); weighted variant
    sum_i CE_i * w_i/||w||_1."""
    logits = np.asarray(logits, F32)
    n, c = logits.shape
    logp = log_softmax(logits)
    t = np.zeros((n, c), F32)
    t[np.arange(n), np.asarray(targets)] = 1
    t = (F32(1 - eps) * t + F32(eps / c)).astype(F32)
    if weights is not None:
        w = np.asarray(weights, F32)
        per = (-t * logp).sum(axis=1, dtype=F32)
        w = w / max(np.abs(w).sum(dtype=F32), F32(1e-12))
        return F32((per * w).sum(dtype=F32))
    return F32((-t * logp).mean(axis=0, dtype=F32).sum(dtype=F32))


def part_pairwise_dist(emb_knd, eps=1e-16):
    """_part_based_pairwise_distance_matrix, part_averaged_triplet_loss.py:77-93. emb [K,N,D] -> [K,N,N]."""
    e = np.asarray(emb_knd, F32)
    dot = np.einsum("knd,kmd->knm", e, e).astype(F32)
    sq = np.diagonal(dot, axis1=1, axis2=2)
    d2 = sq[:, :, None] - F32(2) * dot + sq[:, None, :]
    d2 = np.maximum(d2, F32(0)).astype(F32)
    mask = (d2 == 0).astype(F32)
    d = np.sqrt(d2 + mask * F32(eps), dtype=F32)
    return (d * (F32(1) - mask)).astype(F32)


def masked_mean(x, mask):
    """tensortools.py:10-19 ; -1 where no part could be compared."""
    x = np.asarray(x, F32)
    m = np.asarray(mask).astype(F32)
    w = m.sum(0, dtype=F32)
    wz = w + (w == 0)
    out = (x * m).sum(0, dtype=F32) / wz
    inv = (w == 0)
    return (out * (~inv) + inv * F32(-1)).astype(F32)


def hard_mine(dist_1nn, labels, margin=0.3):
    """_hard_mine_triplet_loss, part_averaged_triplet_loss.py:95-173.  dist [Kd,N,N] (Kd=1 for the averaged loss).
    Returns (loss, trivial_ratio, valid_ratio) or None when no valid triplet exists."""
    d = np.asarray(dist_1nn, F32)
    labels = np.asarray(labels)
    n = labels.shape[0]
    valid = d != F32(-1)
    same = labels[None, :] == labels[:, None]
    pos = (same & ~np.eye(n, dtype=bool))[None] & valid
    neg = (~same)[None] & valid
    vpos = d * pos.astype(F32) - (~pos).astype(F32)
    hp = vpos.max(-1)
    vneg = d * neg.astype(F32) + (~neg).astype(F32) * FLT_MAX
    hn = vneg.min(-1)
    vt = (hp != F32(-1)) & (hn != FLT_MAX)
    if vt.sum() == 0:
        return None
    hp_v, hn_v = hp[vt], hn[vt]
    m = margin if margin > 0 else 0.3
    tl = np.maximum(hp_v - hn_v + F32(m), F32(0)).astype(F32)
    hard = F32(tl.mean(dtype=F32))
    trivial = F32((tl == 0).sum() / tl.size)
    valid_ratio = F32(vt.sum() / vt.size)
    if margin > 0:
        return hard, trivial, valid_ratio
    # soft margin: mean log(1+exp(-(hn-hp)))  (part_averaged_triplet_loss.py:182-195)
    z = (hn_v - hp_v).astype(F32)
    soft = F32(np.log1p(np.exp(-z)).astype(F32).mean(dtype=F32))
    if not np.isfinite(soft):
        return hard, trivial, valid_ratio
    return soft, trivial, valid_ratio


def part_averaged_triplet(emb_nkd, labels, margin=0.3, parts_visibility=None):
    """PartAveragedTripletLoss.forward, part_averaged_triplet_loss.py:35-65."""
    e = np.asarray(emb_nkd, F32).transpose(1, 0, 2)
    pd = part_pairwise_dist(e)
    if parts_visibility is not None:
        v = np.asarray(parts_visibility).T                   # [K, N]
        if v.dtype == np.bool_:
            m = v[:, None, :] & v[:, :, None]
        else:
            m = np.sqrt(v[:, None, :].astype(F32) * v[:, :, None].astype(F32)).astype(F32)
        dist = masked_mean(pd, m)
    else:
        dist = pd.mean(0, dtype=F32)
    return hard_mine(dist[None], labels, margin)


def gilt(emb, vis, ids, pids, weights=None, use_visibility_scores=False, margin=0.3):
    """GiLtLoss.forward GiLt_loss.py:45-93.  Returns (loss, summary dict of floats)."""
    weights = weights or DEFAULT_WEIGHTS
    summary, terms = {}, []
    for key in [GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS]:
        info = summary.setdefault(key, {})
        w = weights[key]["id"]
        if w > 0:
            s, v, p = np.asarray(ids[key], F32), np.asarray(vis[key]), np.asarray(pids)
            if s.ndim == 3:
                m = s.shape[1]
                s = s.reshape(-1, s.shape[-1])
                p = np.repeat(p[:, None], m, 1).reshape(-1)
                v = v.reshape(-1)
            wts = None
            if use_visibility_scores and v.dtype == np.bool_:
                s, p = s[v], p[v]
            elif use_visibility_scores:
                wts = v
            ce = cross_entropy_ls(s, p, 0.1, wts)
            info["c"] = float(ce)
            info["a"] = float((s.argmax(1) == p).mean()) if len(p) else 0.0
            terms.append(F32(w) * ce)
    for key in [GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS]:
        info = summary.setdefault(key, {})
        w = weights[key]["tr"]
        if w > 0:
            e = np.asarray(emb[key], F32)
            e = e if e.ndim == 3 else e[:, None]
            v = None
            if use_visibility_scores:
                v = np.asarray(vis[key])
                v = v if v.ndim == 2 else v[:, None]
            r = part_averaged_triplet(e, pids, margin, v)
            if r is None:
                raise RuntimeError("no valid triplets (reference would crash at GiLt_loss.py:79)")
            info["t"], info["tt"], info["vt"] = float(r[0]), float(r[1]), float(r[2])
            terms.append(F32(w) * r[0])
    if not terms:
        return F32(0), summary
    return F32(np.sum(np.asarray(terms, F32), dtype=F32)), summary


def bilinear_align_corners(x, out_h, out_w):
    """F.interpolate(mode='bilinear', align_corners=True) on [N,C,H,W] float32."""
    x = np.asarray(x, F32)
    n, c, h, w = x.shape

    def axis(inp, out):
        if out == 1:
            src = np.zeros(1, F32)
        else:
            src = (np.arange(out, dtype=F32) * F32((inp - 1) / (out - 1))).astype(F32)
        i0 = np.minimum(np.floor(src).astype(np.int64), inp - 1)
        i1 = np.minimum(i0 + 1, inp - 1)
        return i0, i1, (src - i0.astype(F32)).astype(F32)

    y0, y1, fy = axis(h, out_h)
    x0, x1, fx = axis(w, out_w)
    top = x[:, :, y0][:, :, :, x0] * (1 - fx) + x[:, :, y0][:, :, :, x1] * fx
    bot = x[:, :, y1][:, :, :, x0] * (1 - fx) + x[:, :, y1][:, :, :, x1] * fx
    return (top * (1 - fy)[:, None] + bot * fy[:, None]).astype(F32)


def body_part_attention(pixel_logits, target_masks, label_smoothing=0.1):
    """part_based_engine.py:118-124 + body_part_attention_loss.py:31-52 (nn.CrossEntropyLoss(label_smoothing=0.1),
    mean over N*Hf*Wf pixels).  Returns (loss, accuracy)."""
    lg = np.asarray(pixel_logits, F32)
    n, c, hf, wf = lg.shape
    tm = bilinear_align_corners(target_masks, hf, wf)
    tgt = tm.argmax(1).reshape(-1)
    rows = lg.transpose(0, 2, 3, 1).reshape(-1, c)
    logp = log_softmax(rows)
    nll = -logp[np.arange(rows.shape[0]), tgt]
    smooth = -logp.sum(1, dtype=F32) / F32(c)
    loss = ((1 - label_smoothing) * nll + label_smoothing * smooth).astype(F32).mean(dtype=F32)
    acc = float((rows.argmax(1) == tgt).mean())
    return F32(loss), acc
