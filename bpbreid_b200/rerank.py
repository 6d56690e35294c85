"""k-reciprocal re-ranking on the device: host-side mirror of `torchreid.utils.re_ranking`
(torchreid/utils/rerank.py:31-113; called by ImagePartBasedEngine._evaluate, part_based_engine.py:218-226).

`re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3)` keeps the reference's signature: numpy arrays in,
numpy array out (the engine passes `.numpy()` matrices).  `re_ranking_device` takes and returns CUDA tensors and never
leaves the GPU; both run the C-ABI stages of csrc/rerank.cu (include/bpbreid_sm100.h: bpb_rerank_*).  There is no CPU
path: without the CUDA library the call raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _dev_f32(x, device):
    t = torch.as_tensor(x)
    return t.to(device=device, dtype=torch.float32).contiguous()


def re_ranking_device(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3, return_stages=False):
    """All three matrices are CUDA float tensors ([Nq,Ng], [Nq,Nq], [Ng,Ng]); returns the re-ranked [Nq,Ng] distance
    as a CUDA float32 tensor (rerank.py:111-113)."""
    if not (q_g_dist.is_cuda and q_q_dist.is_cuda and g_g_dist.is_cuda):
        raise RuntimeError("re_ranking_device: the distance matrices must be CUDA tensors (no CPU path)")
    dev = q_g_dist.device
    qg, qq, gg = (_dev_f32(t, dev) for t in (q_g_dist, q_q_dist, g_g_dist))
    nq, ng = qg.shape
    if qq.shape != (nq, nq) or gg.shape != (ng, ng):
        raise ValueError(f"re_ranking: inconsistent shapes q_g {tuple(qg.shape)}, q_q {tuple(qq.shape)}, g_g {tuple(gg.shape)}")
    k1, k2 = int(k1), int(k2)
    if k1 < 1 or k2 < 1:
        raise ValueError(f"re_ranking: k1={k1}, k2={k2} must be >= 1")
    a = nq + ng
    s = _lib.stream_ptr()
    with torch.cuda.device(dev):
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        cmax = torch.empty(a, **f32)
        orig = torch.empty(a, a, **f32)
        _lib.call("bpb_rerank_build", qg.data_ptr(), qq.data_ptr(), gg.data_ptr(), nq, ng, cmax.data_ptr(), orig.data_ptr(), s)
        rk = min(max(k1 + 1, k2), a)
        rank = torch.empty(a, rk, **i32)
        _lib.call("bpb_rerank_topk", orig.data_ptr(), a, rk, rank.data_ptr(), s)
        cap = _lib.query("bpb_rerank_row_capacity", k1)
        vidx, vval, vcnt = torch.empty(a, cap, **i32), torch.empty(a, cap, **f32), torch.empty(a, **i32)
        _lib.call("bpb_rerank_kreciprocal", orig.data_ptr(), rank.data_ptr(), a, rk, k1, vidx.data_ptr(), vval.data_ptr(),
                  vcnt.data_ptr(), s)
        qcap = min(k2, a) * cap
        qidx, qval, qcnt = torch.empty(a, qcap, **i32), torch.empty(a, qcap, **f32), torch.empty(a, **i32)
        _lib.call("bpb_rerank_qexpand", rank.data_ptr(), rk, k2, vidx.data_ptr(), vval.data_ptr(), vcnt.data_ptr(), k1, a,
                  qidx.data_ptr(), qval.data_ptr(), qcnt.data_ptr(), s)
        colcnt = torch.empty(a, **i32)
        _lib.call("bpb_rerank_csc_count", qidx.data_ptr(), qcnt.data_ptr(), qcap, a, colcnt.data_ptr(), s)
        col_ptr = torch.zeros(a + 1, **i32)
        col_ptr[1:] = torch.cumsum(colcnt, 0, dtype=torch.int32)
        nnz = int(col_ptr[-1])                                       # sizes the column-compressed copy (one host read)
        crow, cval, cursor = torch.empty(max(nnz, 1), **i32), torch.empty(max(nnz, 1), **f32), torch.empty(a, **i32)
        _lib.call("bpb_rerank_csc_fill", qidx.data_ptr(), qval.data_ptr(), qcnt.data_ptr(), qcap, a, col_ptr.data_ptr(),
                  cursor.data_ptr(), crow.data_ptr(), cval.data_ptr(), s)
        out = torch.empty(nq, ng, **f32)
        scratch = None
        if a * 4 > 200 * 1024:
            scratch = torch.empty(_lib.query("bpb_rerank_jaccard_blocks", nq, a) * a, **f32)
        lam = np.float32(lambda_value)
        one_minus = np.float32(1 - lambda_value)                    # rerank.py:106: python double, cast once to float32
        _lib.call("bpb_rerank_jaccard", orig.data_ptr(), qidx.data_ptr(), qval.data_ptr(), qcnt.data_ptr(), qcap,
                  col_ptr.data_ptr(), crow.data_ptr(), cval.data_ptr(), nq, a, float(one_minus), float(lam),
                  _lib.ptr(scratch), out.data_ptr(), s)
    if return_stages:
        return out, dict(orig=orig, rank=rank, vidx=vidx, vval=vval, vcnt=vcnt, qidx=qidx, qval=qval, qcnt=qcnt,
                         col_ptr=col_ptr, crow=crow, cval=cval)
    return out


def re_ranking(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3):
    """Drop-in for torchreid.utils.re_ranking (rerank.py:31): numpy (or CPU tensor) matrices in, numpy float32 out; CUDA
    tensors in, CUDA tensor out."""
    if torch.is_tensor(q_g_dist) and q_g_dist.is_cuda:
        return re_ranking_device(q_g_dist, q_q_dist, g_g_dist, k1, k2, lambda_value)
    if not torch.cuda.is_available():
        raise RuntimeError("bpbreid_b200.rerank.re_ranking runs on sm_100a only (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    out = re_ranking_device(_dev_f32(q_g_dist, dev), _dev_f32(q_q_dist, dev), _dev_f32(g_g_dist, dev), k1, k2, lambda_value)
    return out.cpu().numpy()
