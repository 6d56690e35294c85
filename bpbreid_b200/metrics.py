"""Host-side mirror of the reference evaluation API for the BPBReID hot path (sm_100a kernels behind it).

    compute_distance_matrix_using_bp_features(qf, gf, qf_parts_visibility, gf_parts_visibility, dist_combine_strat,
        batch_size_pairwise_dist_matrix, use_gpu, metric) -> (dist [Nq,Ng], part_dist [P,Nq,Ng])  CPU fp32 tensors
                                                                       torchreid/metrics/distance.py:87-99
    evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, eval_metric, ...) -> {'cmc','mAP'}
                                                                       torchreid/metrics/rank.py:173-214
    evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03) -> (cmc, mAP)
                                                                       torchreid/metrics/rank_cylib/rank_cy.pyx:24-32

Host buffers in, host buffers out -- exactly the reference's contract (part_based_engine.py:211-217 calls
.numpy() on both distance outputs).  `part_distance_device` / `rank_device` are the device-resident variants the
benchmark times for the "inputs already in HBM" number.  No CPU fallback: a machine without the CUDA library raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

_METRICS = {"euclidean": 0, "cosine": 1}
_STRATS = {"mean": 0, "max": 1}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("bpbreid_b200.metrics needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def part_distance_device(qf, gf, qv=None, gv=None, dist_combine_strat="mean", metric="euclidean", impl=0,
                         want_part_dist=True, out=None):
    """Device tensors in -> device tensors out.  qf [Nq,P,D], gf [Ng,P,D] fp32; qv/gv [N,P] bool or float."""
    if metric not in _METRICS:
        raise ValueError('Unknown distance metric: {}. Please choose either "euclidean" or "cosine"'.format(metric))
    if dist_combine_strat not in _STRATS:
        raise ValueError('Body parts distance combination strategy "{}" not supported'.format(dist_combine_strat))
    qf = qf.float().contiguous()
    gf = gf.float().contiguous()
    nq, p, d = qf.shape
    ng = gf.shape[0]
    vis_mode, qvt, gvt = 0, None, None
    if qv is not None and gv is not None:
        if qv.dtype is torch.bool and gv.dtype is torch.bool:
            vis_mode, qvt, gvt = 1, qv.contiguous().view(torch.uint8), gv.contiguous().view(torch.uint8)
        else:
            vis_mode, qvt, gvt = 2, qv.float().contiguous(), gv.float().contiguous()
    dev = qf.device
    if out is None:
        dist = torch.empty(nq, ng, device=dev, dtype=torch.float32)
        part = torch.empty(p, nq, ng, device=dev, dtype=torch.float32) if want_part_dist else None
        scratch = torch.empty(_lib.query("bpb_part_distance_scratch_floats", nq, ng, p), device=dev)
    else:
        dist, part, scratch = out
    _lib.call("bpb_part_distance", qf.data_ptr(), gf.data_ptr(), _lib.ptr(qvt), _lib.ptr(gvt), vis_mode, nq, ng, p,
              d, _METRICS[metric], _STRATS[dist_combine_strat], int(impl), dist.data_ptr(), _lib.ptr(part),
              scratch.data_ptr(), _lib.stream_ptr())
    return dist, part


def compute_distance_matrix_using_bp_features(qf, gf, qf_parts_visibility=None, gf_parts_visibility=None,
                                              dist_combine_strat="mean", batch_size_pairwise_dist_matrix=5000,
                                              use_gpu=False, metric="euclidean", exact=False):
    """Reference signature (+ `exact`: the SIMT fp32 distance path instead of the tcgen05 tf32 one, for evaluations that
    must be rank-for-rank comparable with the fp32 reference).  `batch_size_pairwise_dist_matrix` only bounded the reference's memory (gallery
    chunks + a D2H copy per chunk); one fused launch covers the whole matrix here, so it is accepted and ignored.
    `use_gpu` is accepted for compatibility: the computation always runs on the GPU."""
    dev = _device()
    q = qf.to(dev, non_blocking=True)
    g = gf.to(dev, non_blocking=True)
    qv = None if qf_parts_visibility is None else qf_parts_visibility.to(dev, non_blocking=True)
    gv = None if gf_parts_visibility is None else gf_parts_visibility.to(dev, non_blocking=True)
    dist, part = part_distance_device(q, g, qv, gv, dist_combine_strat, metric, impl=1 if exact else 0)
    return dist.cpu(), part.cpu()


def rank_device(dist, q_pids, g_pids, q_camids, g_camids, max_rank=50):
    """Device-resident market1501 evaluation. Returns a dict of device tensors (no sync)."""
    dist = dist.float().contiguous()
    nq, ng = dist.shape
    dev = dist.device
    mr = min(max_rank, ng)
    i64 = lambda a: torch.as_tensor(a, dtype=torch.int64).to(dev).contiguous()  # noqa: E731
    qp, gp, qc, gc = i64(q_pids), i64(g_pids), i64(q_camids), i64(g_camids)
    ap64 = torch.empty(nq, device=dev, dtype=torch.float64)
    ap32 = torch.empty(nq, device=dev, dtype=torch.float32)
    first = torch.empty(nq, device=dev, dtype=torch.int32)
    cmc = torch.empty(mr, device=dev, dtype=torch.float32)
    map64 = torch.empty(1, device=dev, dtype=torch.float64)
    nvalid = torch.empty(1, device=dev, dtype=torch.int64)
    map32 = torch.empty(1, device=dev, dtype=torch.float32)
    overflow = torch.empty(1, device=dev, dtype=torch.int32)
    _lib.call("bpb_rank_market1501", dist.data_ptr(), qp.data_ptr(), gp.data_ptr(), qc.data_ptr(), gc.data_ptr(),
              nq, ng, mr, ap64.data_ptr(), ap32.data_ptr(), first.data_ptr(), cmc.data_ptr(), map64.data_ptr(),
              nvalid.data_ptr(), map32.data_ptr(), overflow.data_ptr(), _lib.stream_ptr())
    return dict(cmc=cmc, mAP=map64, mAP32=map32, num_valid=nvalid, ap64=ap64, ap32=ap32, first_rank=first,
                overflow=overflow)


def argsort_rows(dist):
    """Stable ascending argsort of each row on the device (== np.argsort(kind='stable')); returns int32 [Nq,Ng]."""
    dist = dist.float().contiguous()
    nq, ng = dist.shape
    idx = torch.empty(nq, ng, device=dist.device, dtype=torch.int32)
    _lib.call("bpb_argsort_rows", dist.data_ptr(), nq, ng, idx.data_ptr(), _lib.stream_ptr())
    return idx


def _check_rank(res, ng, max_rank):
    if int(res["overflow"].item()) != 0:
        raise RuntimeError("rank: a query has more than 2048 same-identity gallery entries")
    assert int(res["num_valid"].item()) > 0, "Error: all query identities do not appear in gallery"
    if ng < max_rank:
        print("Note: number of gallery samples is quite small, got {}".format(ng))


def evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50, eval_metric="default", q_anns=None,
                  g_anns=None, use_cython=True):
    """Reference signature (rank.py:173).  numpy in, {'cmc': ndarray[max_rank], 'mAP': float} out."""
    if eval_metric == "cuhk03":
        raise NotImplementedError("cuhk03 single-gallery-shot protocol (10 random repeats, nondeterministic) is "
                                  "outside the hot path (SURVEY.md section 8a row a16)")
    if eval_metric != "default":
        raise ValueError("Incorrect eval_metric value '{}'".format(eval_metric))
    dev = _device()
    d = torch.as_tensor(np.ascontiguousarray(distmat, dtype=np.float32)).to(dev)
    res = rank_device(d, q_pids, g_pids, q_camids, g_camids, max_rank)
    _check_rank(res, d.shape[1], max_rank)
    return {"cmc": res["cmc"].cpu().numpy(), "mAP": float(res["mAP"].item())}


def evaluate_cy(distmat, q_pids, g_pids, q_camids, g_camids, max_rank, use_metric_cuhk03=False):
    """Reference's native entry point (rank_cy.pyx:24): returns the (cmc, mAP) tuple with fp32 accumulation."""
    if use_metric_cuhk03:
        raise NotImplementedError("cuhk03 protocol is outside the hot path")
    dev = _device()
    d = torch.as_tensor(np.ascontiguousarray(distmat, dtype=np.float32)).to(dev)
    res = rank_device(d, q_pids, g_pids, q_camids, g_camids, max_rank)
    _check_rank(res, d.shape[1], max_rank)
    return res["cmc"].cpu().numpy().astype(np.float32), float(res["mAP32"].item())
