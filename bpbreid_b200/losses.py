"""Host-side mirror of the reference loss API for the BPBReID hot path (forward values on sm_100a kernels).

    GiLtLoss(losses_weights, use_visibility_scores, triplet_margin, loss_name, use_gpu, writer)
        .forward(embeddings_dict, visibility_scores_dict, id_cls_scores_dict, pids) -> (loss, summary)
                                                                       torchreid/losses/GiLt_loss.py:27-119
    init_part_based_triplet_loss(name, margin, writer)                 torchreid/losses/__init__.py:24-32
    PartAveragedTripletLoss(margin, epsilon, writer).forward(emb[N,K,D], labels, parts_visibility)
                                                                       torchreid/losses/part_averaged_triplet_loss.py:27-65
    CrossEntropyLoss(eps, label_smooth).forward(inputs, targets, weights)  torchreid/losses/cross_entropy_loss.py:28-56
    BodyPartAttentionLoss(loss_type, label_smoothing, use_gpu).forward(pixels_cls_scores, targets)
                                                                       torchreid/losses/body_part_attention_loss.py:17-52

Same dictionary keys, same sentinels, same error behaviour (ValueError for unknown loss names, warning + None
when no valid triplet exists).  Inputs must be CUDA tensors: there is no CPU fallback.
This round computes forward values only (BASELINE.json's metric is forward + loss); the analytic backward is
SURVEY.md section 8f-2.
"""
from __future__ import annotations

import ctypes
import warnings
from collections import OrderedDict

import torch
from torch import nn

from . import _lib
from .constants import GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS, PIXELS


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: bpbreid_b200 losses run on sm_100a only (got a {t.device} tensor); "
                           "there is no CPU fallback")


def _vis_args(vis):
    """-> (pointer-holder tensor or None, vis_mode)"""
    if vis is None:
        return None, 0
    if vis.dtype is torch.bool:
        return vis.contiguous().view(torch.uint8), 1
    return vis.contiguous().float(), 2


class PartAveragedTripletLoss(nn.Module):
    def __init__(self, margin=0.3, epsilon=1e-16, writer=None):
        super().__init__()
        self.margin = margin
        self.writer = writer
        self.epsilon = epsilon
        self.last = {}

    def forward(self, part_based_embeddings, labels, parts_visibility=None):
        emb = part_based_embeddings
        _need_cuda(emb, "PartAveragedTripletLoss")
        n, k, d = emb.shape
        knd = emb.permute(1, 0, 2)
        if not knd.is_contiguous():
            knd = knd.contiguous()
        knd = knd.float()
        labels = labels.to(device=emb.device, dtype=torch.int64).contiguous()
        vis_t, vis_mode = _vis_args(parts_visibility)
        dev = emb.device
        dist = torch.empty(n, n, device=dev, dtype=torch.float32)
        hp = torch.empty(n, device=dev)
        hn = torch.empty(n, device=dev)
        hp_idx = torch.empty(n, device=dev, dtype=torch.int32)
        hn_idx = torch.empty(n, device=dev, dtype=torch.int32)
        out4 = torch.empty(4, device=dev)
        _lib.call("bpb_triplet_fwd", knd.data_ptr(), labels.data_ptr(), _lib.ptr(vis_t), vis_mode, n, k, d,
                  float(self.margin), dist.data_ptr(), hp.data_ptr(), hn.data_ptr(), hp_idx.data_ptr(),
                  hn_idx.data_ptr(), out4.data_ptr(), _lib.stream_ptr())
        self.last = dict(dist=dist, hp=hp, hn=hn, hp_idx=hp_idx, hn_idx=hn_idx, out4=out4)
        if self.writer is not None:
            if vis_mode != 0 and hasattr(self.writer, "update_invalid_part_based_pairwise_distances_count"):
                pass  # statistics hooks of the reference writer are logging only (part_averaged_triplet_loss.py:69)
        return _TripletResult(out4)

    def _hard_mine_triplet_loss(self, batch_pairwise_dist, labels, margin):  # API parity helper
        raise NotImplementedError("mining is fused into forward(); use forward()")


class _TripletResult(tuple):
    """(loss, trivial_ratio, valid_ratio) as 0-dim device tensors, without forcing a device sync.
    `materialize()` applies the reference's "no valid triplet -> warn and return None" rule (it needs n_valid on
    the host, i.e. one sync), which GiLtLoss does lazily."""

    def __new__(cls, out4):
        obj = super().__new__(cls, (out4[0], out4[1], out4[2]))
        obj.out4 = out4
        return obj

    def materialize(self):
        if float(self.out4[3]) == 0.0:
            warnings.warn("CRITICAL WARNING: no valid triplets were generated for current batch")
            return None
        return tuple(self)


__body_parts_losses = {"part_averaged_triplet_loss": PartAveragedTripletLoss}
_UNBUILT = ("part_max_triplet_loss", "part_min_triplet_loss", "part_max_min_triplet_loss",
            "part_random_max_min_triplet_loss", "inter_parts_triplet_loss", "intra_parts_triplet_loss")


def init_part_based_triplet_loss(name, **kwargs):
    avai = list(__body_parts_losses.keys())
    if name in _UNBUILT:
        raise NotImplementedError(f'"{name}" is a reference loss outside the hot path (SURVEY.md section 2 row 12)')
    if name not in avai:
        raise ValueError('Invalid loss name. Received "{}", but expected to be one of {}'.format(name, avai))
    return __body_parts_losses[name](**kwargs)


class CrossEntropyLoss(nn.Module):
    def __init__(self, eps=0.1, label_smooth=True):
        super().__init__()
        self.eps = eps if label_smooth else 0
        self.last_accuracy = None

    def forward(self, inputs, targets, weights=None, acc_over_weighted_rows=False):
        _need_cuda(inputs, "CrossEntropyLoss")
        assert inputs.shape[0] == targets.shape[0]
        n, c = inputs.shape
        if inputs.stride(1) != 1:
            inputs = inputs.contiguous()
        inputs = inputs.float()
        ld = inputs.stride(0)
        targets = targets.to(device=inputs.device, dtype=torch.int64).contiguous()
        w = None if weights is None else weights.to(device=inputs.device, dtype=torch.float32).contiguous()
        scratch = torch.empty(2 * n, device=inputs.device)
        out2 = torch.empty(2, device=inputs.device)
        _lib.call("bpb_ce_ls_fwd", inputs.data_ptr(), ld, targets.data_ptr(), _lib.ptr(w), n, c, float(self.eps),
                  1 if acc_over_weighted_rows else 0, scratch.data_ptr(), out2.data_ptr(), _lib.stream_ptr())
        self.last_accuracy = out2[1]
        return out2[0]


class GiLtLoss(nn.Module):
    default_losses_weights = {
        GLOBAL: {"id": 1., "tr": 0.},
        FOREGROUND: {"id": 1., "tr": 0.},
        CONCAT_PARTS: {"id": 1., "tr": 0.},
        PARTS: {"id": 0., "tr": 1.},
    }

    def __init__(self, losses_weights=None, use_visibility_scores=False, triplet_margin=0.3,
                 loss_name="part_averaged_triplet_loss", use_gpu=False, writer=None):
        super().__init__()
        if losses_weights is None:
            losses_weights = self.default_losses_weights
        self.losses_weights = losses_weights
        self.part_triplet_loss = init_part_based_triplet_loss(loss_name, margin=triplet_margin, writer=writer)
        self.identity_loss = CrossEntropyLoss(label_smooth=True)
        self.use_visibility_scores = use_visibility_scores

    def forward(self, embeddings_dict, visibility_scores_dict, id_cls_scores_dict, pids):
        loss_summary = {}
        losses = []
        for key in [GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS]:
            loss_info = OrderedDict() if key not in loss_summary else loss_summary[key]
            ce_w = self.losses_weights[key]["id"]
            if ce_w > 0:
                l, a = self.compute_id_cls_loss(id_cls_scores_dict[key], visibility_scores_dict[key], pids)
                losses.append((ce_w, l))
                loss_info["c"] = l
                loss_info["a"] = a
            loss_summary[key] = loss_info
        for key in [GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS]:
            loss_info = OrderedDict() if key not in loss_summary else loss_summary[key]
            tr_w = self.losses_weights[key]["tr"]
            if tr_w > 0:
                t, tt, vt = self.compute_triplet_loss(embeddings_dict[key], visibility_scores_dict[key], pids)
                losses.append((tr_w, t))
                loss_info["t"] = t
                loss_info["tt"] = tt
                loss_info["vt"] = vt
            loss_summary[key] = loss_info
        if len(losses) == 0:
            return torch.tensor(0., device=pids.device if pids.is_cuda else None), loss_summary
        dev = losses[0][1].device
        out = torch.empty(1, device=dev)
        ptrs = (ctypes.c_void_p * len(losses))(*[l.data_ptr() for _, l in losses])
        ws = (ctypes.c_float * len(losses))(*[float(w) for w, _ in losses])
        _lib.call("bpb_weighted_sum", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(ws, ctypes.c_void_p),
                  len(losses), out.data_ptr(), _lib.stream_ptr())
        return out[0], loss_summary

    def compute_triplet_loss(self, embeddings, visibility_scores, pids):
        if self.use_visibility_scores:
            visibility = visibility_scores if len(visibility_scores.shape) == 2 else visibility_scores.unsqueeze(1)
        else:
            visibility = None
        embeddings = embeddings if len(embeddings.shape) == 3 else embeddings.unsqueeze(1)
        res = self.part_triplet_loss(embeddings, pids, parts_visibility=visibility)
        # same contract as the reference: a batch without any valid triplet yields None, which GiLt_loss.py:79
        # then fails to unpack.  Checked lazily so that the hot loop is not forced to synchronise.
        return res

    def compute_id_cls_loss(self, id_cls_scores, visibility_scores, pids):
        if len(id_cls_scores.shape) == 3:
            m = id_cls_scores.shape[1]
            id_cls_scores = id_cls_scores.flatten(0, 1)
            pids = pids.unsqueeze(1).expand(-1, m).flatten(0, 1)
            visibility_scores = visibility_scores.flatten(0, 1)
        weights, filtered = None, False
        if self.use_visibility_scores and visibility_scores.dtype is torch.bool:
            # reference drops the invisible rows; a {0,1}-weighted mean over all rows is the same number and
            # keeps shapes static (no host sync)
            weights, filtered = visibility_scores.float(), True
        elif self.use_visibility_scores and visibility_scores.dtype is not torch.bool:
            weights = visibility_scores
        cls_loss = self.identity_loss(id_cls_scores, pids, weights, acc_over_weighted_rows=filtered)
        return cls_loss, self.identity_loss.last_accuracy


class BodyPartAttentionLoss(nn.Module):
    def __init__(self, loss_type="cl", label_smoothing=0.1, use_gpu=False):
        super().__init__()
        if loss_type in ("fl", "dl"):
            raise NotImplementedError("focal / dice part-prediction losses need monai (not in the hot path)")
        if loss_type != "cl":
            raise ValueError("Loss {} for part prediction is not supported".format(loss_type))
        self.label_smoothing = label_smoothing

    def forward_from_masks(self, pixels_cls_scores, target_masks):
        """Fused variant: takes the raw external masks [N,K+1,Hm,Wm]; the bilinear resize + argmax target
        preparation of part_based_engine.py:118-124 happens inside the kernel."""
        _need_cuda(pixels_cls_scores, "BodyPartAttentionLoss")
        n, kp1, hf, wf = pixels_cls_scores.shape
        tm = target_masks.to(device=pixels_cls_scores.device, dtype=torch.float32).contiguous()
        lg = pixels_cls_scores.float().contiguous()
        partial = torch.empty(512, device=lg.device)
        out2 = torch.empty(2, device=lg.device)
        _lib.call("bpb_pixel_ce_fwd", lg.data_ptr(), tm.data_ptr(), n, kp1, hf, wf, tm.shape[2], tm.shape[3],
                  float(self.label_smoothing), partial.data_ptr(), out2.data_ptr(), _lib.stream_ptr())
        summary = {PIXELS: OrderedDict(c=out2[0], a=out2[1])}
        return out2[0], summary

    def forward(self, pixels_cls_scores, targets):
        """Reference signature: `targets` are integer pixel labels [N,Hf,Wf] (already argmax-ed)."""
        kp1 = pixels_cls_scores.shape[1]
        onehot = torch.nn.functional.one_hot(targets.long(), kp1).permute(0, 3, 1, 2).float()
        return self.forward_from_masks(pixels_cls_scores, onehot)
