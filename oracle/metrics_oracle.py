"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the evaluation side of the hot path:

  * part-based query x gallery distance -> /root/reference/torchreid/metrics/distance.py:87-247
    (masked_mean / replace_values       -> /root/reference/torchreid/utils/tensortools.py:1-19)
  * CMC / mAP ranking (market1501)      -> /root/reference/torchreid/metrics/rank.py:97-159
    native twin                         -> /root/reference/torchreid/metrics/rank_cylib/rank_cy.pyx:154-249
    (C restatement in oracle/rank_oracle.c, loaded below through ctypes when it has been built)
  * L2 normalisation + test-embedding assembly -> /root/reference/torchreid/engine/engine.py:558-559,
    /root/reference/torchreid/engine/image/part_based_engine.py:365-387

Ranking contract (SURVEY.md section 7 hard part 5): ``np.argsort(kind='stable')`` on the identical fp32 distance
matrix, i.e. ties broken by ascending gallery index.  The reference's default argsort is unstable, so it
agrees with this up to permutations inside tie groups; cmc/mAP computed below use the stable order.
Parity pin: tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

F32 = np.float32


def l2_normalize(x, eps=1e-12):
    """F.normalize(p=2, dim=-1) (engine.py:558-559)."""
    x = np.asarray(x, F32)
    n = np.sqrt((x * x).sum(-1, keepdims=True, dtype=F32))
    return (x / np.maximum(n, F32(eps))).astype(F32)


def body_parts_dist(qf, gf, metric="euclidean"):
    """_compute_body_parts_dist_matrices distance.py:222-247. qf [Nq,P,D], gf [Ng,P,D] -> [P,Nq,Ng]."""
    q = np.asarray(qf, F32).transpose(1, 0, 2)
    g = np.asarray(gf, F32).transpose(1, 0, 2)
    dot = np.matmul(q, g.transpose(0, 2, 1)).astype(F32)
    if metric == "euclidean":
        qs = (q * q).sum(-1, dtype=F32)
        gs = (g * g).sum(-1, dtype=F32)
        d = qs[:, :, None] - F32(2) * dot + gs[:, None, :]
        return np.sqrt(np.maximum(d, F32(0)), dtype=F32)
    if metric == "cosine":
        return (F32(1) - dot).astype(F32)
    raise ValueError(f'Unknown distance metric: {metric}. Please choose either "euclidean" or "cosine"')


def part_distance(qf, gf, qv=None, gv=None, dist_combine_strat="mean", metric="euclidean"):
    """compute_distance_matrix_using_bp_features distance.py:87-219 (all three visibility variants).
    Returns (dist [Nq,Ng], part_dist [P,Nq,Ng]) float32.  Gallery chunking in the reference only bounds
    memory; the result is chunk-independent so it is not restated."""
    pd = body_parts_dist(qf, gf, metric)
    if qv is None or gv is None:
        if dist_combine_strat == "max":
            return pd.max(0), pd
        if dist_combine_strat == "mean":
            return pd.mean(0, dtype=F32), pd
        raise ValueError(f'Body parts distance combination strategy "{dist_combine_strat}" not supported')
    qv, gv = np.asarray(qv), np.asarray(gv)
    if qv.dtype == np.bool_ and gv.dtype == np.bool_:
        m = qv.T[:, :, None] & gv.T[:, None, :]
        mf = m.astype(F32)
        valid_pd = (pd * mf + (1 - mf) * F32(-1)).astype(F32)
        if dist_combine_strat == "max":
            dist = valid_pd.max(0)
        elif dist_combine_strat == "mean":
            w = mf.sum(0, dtype=F32)
            dist = (pd * mf).sum(0, dtype=F32) / (w + (w == 0))
            dist = np.where(w == 0, F32(-1), dist).astype(F32)
        else:
            raise ValueError(f'Body parts distance combination strategy "{dist_combine_strat}" not supported')
        max_value = F32(valid_pd.max() + F32(1))
        dist = np.where(dist == F32(-1), max_value, dist).astype(F32)
        valid_pd = np.where(valid_pd == F32(-1), max_value, valid_pd).astype(F32)
        return dist, valid_pd
    # continuous visibility: weights sqrt(vq*vg) (distance.py:198), per-part matrix returned unmasked
    m = np.sqrt(qv.T.astype(F32)[:, :, None] * gv.T.astype(F32)[:, None, :]).astype(F32)
    w = m.sum(0, dtype=F32)
    dist = (pd * m).sum(0, dtype=F32) / (w + (w == 0))
    dist = np.where(w == 0, F32(-1), dist).astype(F32)
    max_value = F32(pd.max() + F32(1))
    dist = np.where(dist == F32(-1), max_value, dist).astype(F32)
    return dist, pd


def eval_market1501(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, return_all=False):
    """rank.py:97-159 with a stable argsort.  float64 accumulation like numpy in the reference."""
    distmat = np.asarray(distmat, F32)
    q_pids, g_pids = np.asarray(q_pids, np.int64), np.asarray(g_pids, np.int64)
    q_camids, g_camids = np.asarray(q_camids, np.int64), np.asarray(g_camids, np.int64)
    num_q, num_g = distmat.shape
    max_rank = min(max_rank, num_g)
    indices = np.argsort(distmat, axis=1, kind="stable")
    all_cmc, all_ap = [], []
    for qi in range(num_q):
        order = indices[qi]
        keep = ~((g_pids[order] == q_pids[qi]) & (g_camids[order] == q_camids[qi]))
        raw = (g_pids[order] == q_pids[qi])[keep].astype(np.int32)
        if not raw.any():
            continue
        cs = raw.cumsum()
        all_cmc.append(np.minimum(cs, 1)[:max_rank])
        prec = cs / (np.arange(cs.size) + 1.0)
        all_ap.append((prec * raw).sum() / raw.sum())
    assert len(all_ap) > 0, "Error: all query identities do not appear in gallery"
    cmc = np.asarray(all_cmc).astype(F32).sum(0) / float(len(all_ap))
    out = {"cmc": cmc, "mAP": float(np.mean(all_ap))}
    if return_all:
        out["all_AP"] = np.asarray(all_ap)
        out["indices"] = indices
    return out


# ---------------------------------------------------------------------------------------------
# C restatement of rank_cy.pyx (built by oracle/build.py into oracle/_build/librank_oracle.so)
# ---------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = None


def c_lib():
    global _C_LIB
    if _C_LIB is None:
        path = os.path.join(_HERE, "_build", "librank_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle C library not built: run `python oracle/build.py`")
        lib = ctypes.CDLL(path)
        lib.oracle_eval_market1501.restype = ctypes.c_int
        lib.oracle_eval_market1501.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_long] * 3 + [ctypes.c_void_p] * 3
        _C_LIB = lib
    return _C_LIB


def eval_market1501_c(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50):
    """C twin of rank_cy.eval_market1501_cy (float32 accumulators, like the Cython `float`s). -> (cmc, mAP)."""
    lib = c_lib()
    d = np.ascontiguousarray(distmat, F32)
    qp, gp = np.ascontiguousarray(q_pids, np.int64), np.ascontiguousarray(g_pids, np.int64)
    qc, gc = np.ascontiguousarray(q_camids, np.int64), np.ascontiguousarray(g_camids, np.int64)
    nq, ng = d.shape
    mr = min(max_rank, ng)
    cmc = np.zeros(mr, F32)
    m_ap = np.zeros(1, F32)
    nvalid = np.zeros(1, np.int64)
    rc = lib.oracle_eval_market1501(d.ctypes.data, qp.ctypes.data, gp.ctypes.data, qc.ctypes.data, gc.ctypes.data,
                                    nq, ng, mr, cmc.ctypes.data, m_ap.ctypes.data, nvalid.ctypes.data)
    if rc != 0:
        raise AssertionError("Error: all query identities do not appear in gallery")
    return cmc, float(m_ap[0])
