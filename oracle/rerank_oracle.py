"""ORACLE (test infrastructure only; never imported by the product path).

CPU restatement of the reference's k-reciprocal re-ranking, `torchreid/utils/rerank.py:31-113` (Zhong et al., CVPR 2017),
as called by `ImagePartBasedEngine._evaluate` (`part_based_engine.py:218-226`) on the query-gallery, query-query and
gallery-gallery part-based distance matrices.  Written stage by stage in the order the CUDA path computes it, so the
tests can compare the intermediate integer results (neighbour table, expanded k-reciprocal sets) bit for bit and the
floating-point ones (weights, Jaccard distance, final distance) within the tolerance written in the test.

Pinned against the unmodified reference: `oracle/make_golden.py rerank` runs `torchreid/utils/rerank.py` (numpy only,
imported from where it lies) on the seeded cases of `tests/golden_inputs.py::rerank_case` and stores its output in
`tests/golden/rerank.npz`; `tests/test_oracle_vs_golden.py` checks this file against it.

Tie handling: the reference ranks with `np.argsort` (introsort, free to permute equal keys, rerank.py:47).  Like the
ranking contract of the evaluation path, the restatement and the CUDA path use the stable order (value, then index).
"""
from __future__ import annotations

import numpy as np


def normalised_distance(q_g, q_q, g_g):
    """rerank.py:36-46: squared distances of the (Nq+Ng)^2 block matrix, every column divided by its maximum, then
    transposed: orig[i, j] = D[j, i]^2 / max_k D[k, i]^2  (float32 throughout)."""
    d = np.concatenate([np.concatenate([q_q, q_g], axis=1), np.concatenate([q_g.T, g_g], axis=1)], axis=0)
    d = np.power(d, 2).astype(np.float32)
    return np.transpose(1. * d / np.max(d, axis=0)).astype(np.float32)


def neighbour_table(orig, k):
    """rerank.py:47,56-57: the first k entries of every row of argsort(orig) (stable order)."""
    return np.argsort(orig, axis=1, kind="stable")[:, :k].astype(np.int32)


def _k_reciprocal(rank, i, k):
    """rerank.py:56-59 / 63-70: members f of the first k neighbours of i that have i among their own first k."""
    fwd = rank[i, :k]
    back = rank[fwd, :k]
    return fwd[np.where(back == i)[0]]


def expanded_sets(rank, k1):
    """rerank.py:54-77: R(i) plus every R_half(c), c in R(i), that shares more than 2/3 of its members with R(i);
    returned as sorted unique index arrays (np.unique)."""
    a = rank.shape[0]
    half = int(np.around(k1 / 2.)) + 1
    out = []
    for i in range(a):
        r = _k_reciprocal(rank, i, k1 + 1)
        exp = r
        for c in r:
            rc = _k_reciprocal(rank, c, half)
            if len(np.intersect1d(rc, r)) > 2. / 3 * len(rc):
                exp = np.append(exp, rc)
        out.append(np.unique(exp))
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, return_stages=False):
    """rerank.py:31-113.  Returns final_dist [Nq, Ng] float32."""
    nq, ng = q_g_dist.shape
    a = nq + ng
    orig = normalised_distance(q_g_dist, q_q_dist, g_g_dist)
    rank = neighbour_table(orig, min(max(k1 + 1, k2), a))
    sets = expanded_sets(rank, k1)
    v = np.zeros((a, a), np.float32)
    for i, e in enumerate(sets):                                   # rerank.py:78-79
        w = np.exp(-orig[i, e])
        v[i, e] = 1. * w / np.sum(w)
    if k2 != 1:                                                    # rerank.py:81-86 (local query expansion)
        v = np.stack([np.mean(v[rank[i, :k2], :], axis=0) for i in range(a)]).astype(np.float32)
    inv = [np.where(v[:, c] != 0)[0] for c in range(a)]            # rerank.py:88-90
    jac = np.zeros((nq, a), np.float32)
    for i in range(nq):                                            # rerank.py:94-104
        t = np.zeros(a, np.float32)
        for c in np.where(v[i, :] != 0)[0]:
            t[inv[c]] = t[inv[c]] + np.minimum(v[i, c], v[inv[c], c])
        jac[i] = 1 - t / (2. - t)
    final = (jac * (1 - lambda_value) + orig[:nq] * lambda_value)[:, nq:].astype(np.float32)   # rerank.py:106-111
    if return_stages:
        return final, dict(orig=orig, rank=rank, sets=sets, v=v, jaccard=jac)
    return final
