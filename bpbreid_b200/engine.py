"""Host-side mirror of the reference engine's methods that sit ON the hot path (everything else of the engine --
data managers, writers, schedulers, checkpoints -- is the reference's own control plane and stays there).

    ImagePartBasedEngine.extract_test_embeddings(model_output)   torchreid/engine/image/part_based_engine.py:365-387
    Engine.normalize(features)                                    torchreid/engine/engine.py:558-559
    ImagePartBasedEngine._feature_extraction(data_loader)         part_based_engine.py:132-166
    ImagePartBasedEngine._evaluate(...) distance + ranking core   part_based_engine.py:186-248
    ImagePartBasedEngine.combine_losses / forward_backward        part_based_engine.py:77-130

Differences that are the point of this framework: features stay ON THE DEVICE between the model, the distance kernel
and the ranking kernel (the reference copies every batch to the host, part_based_engine.py:151-156, and the distance
matrices back again, :216-217); `_evaluate` still returns host numbers (cmc ndarray, mAP float) like the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, metrics
from .constants import PARTS, PIXELS, bn_correspondants


def extract_test_embeddings(model_output, test_embeddings):
    """part_based_engine.py:365-387: select + concatenate the test embeddings -> ([N,P,D], [N,P], [N,P,Hf,Wf], pix)."""
    embeddings, visibility_scores, id_cls_scores, pixels_cls_scores, spatial_features, parts_masks = model_output
    embeddings_list, visibility_scores_list, embeddings_masks_list = [], [], []
    for test_emb in test_embeddings:
        embds = embeddings[test_emb]
        embeddings_list.append(embds if len(embds.shape) == 3 else embds.unsqueeze(1))
        if test_emb in bn_correspondants:
            test_emb = bn_correspondants[test_emb]
        vis_scores = visibility_scores[test_emb]
        visibility_scores_list.append(vis_scores if len(vis_scores.shape) == 2 else vis_scores.unsqueeze(1))
        pt_masks = parts_masks[test_emb]
        embeddings_masks_list.append(pt_masks if len(pt_masks.shape) == 4 else pt_masks.unsqueeze(1))
    assert len(embeddings) != 0
    # torch.cat is a copy (no arithmetic): the outputs no longer alias the forward plan's buffers
    return (torch.cat(embeddings_list, dim=1), torch.cat(visibility_scores_list, dim=1),
            torch.cat(embeddings_masks_list, dim=1), pixels_cls_scores)


def normalize(features, out=None):
    """Engine.normalize: L2-normalise the last dimension on the device (bpb_l2_normalize)."""
    if not features.is_cuda:
        raise RuntimeError("bpbreid_b200.engine.normalize runs on sm_100a only (got a CPU tensor)")
    x = features.float().contiguous()
    y = torch.empty_like(x) if out is None else out
    d = x.shape[-1]
    _lib.call("bpb_l2_normalize", x.data_ptr(), y.data_ptr(), x.numel() // d, d, 1e-12, _lib.stream_ptr())
    return y


class ImagePartBasedEngine:
    """The hot-path subset of torchreid.engine.image.part_based_engine.ImagePartBasedEngine.  Same constructor
    argument names for what it uses; `datamanager`, `writer`, `engine_state`, `scheduler` are accepted and kept so a
    reference script can pass them through unchanged."""

    def __init__(self, datamanager=None, model=None, optimizer=None, writer=None, loss_name="part_averaged_triplet_loss",
                 config=None, dist_combine_strat="mean", batch_size_pairwise_dist_matrix=5000, engine_state=None,
                 margin=0.3, scheduler=None, use_gpu=True, save_model_flag=False, mask_filtering_training=False,
                 mask_filtering_testing=False, losses_weights=None, bpa_loss_type="cl"):
        from . import losses
        self.datamanager, self.model, self.optimizer, self.writer = datamanager, model, optimizer, writer
        self.config, self.engine_state, self.scheduler, self.use_gpu = config, engine_state, scheduler, use_gpu
        self.dist_combine_strat = dist_combine_strat
        self.batch_size_pairwise_dist_matrix = batch_size_pairwise_dist_matrix
        self.mask_filtering_training = mask_filtering_training
        self.mask_filtering_testing = mask_filtering_testing
        cfg_b = config.model.bpbreid if config is not None else model.model_cfg
        self.parts_num = cfg_b.masks.parts_num
        self.test_embeddings = list(cfg_b.test_embeddings)
        if losses_weights is None and config is not None and "loss" in config:
            losses_weights = config.loss.part_based.weights
            bpa_loss_type = config.loss.part_based.ppl
        self.losses_weights = losses_weights
        self.GiLt = losses.GiLtLoss(losses_weights, use_visibility_scores=mask_filtering_training,
                                    triplet_margin=margin, loss_name=loss_name, writer=writer, use_gpu=use_gpu)
        self.body_part_attention_loss = losses.BodyPartAttentionLoss(loss_type=bpa_loss_type, use_gpu=use_gpu)

    # ------------------------------------------------------------------ evaluation
    def extract_test_embeddings(self, model_output):
        return extract_test_embeddings(model_output, self.test_embeddings)

    def normalize(self, features):
        return normalize(features)

    @torch.no_grad()
    def _feature_extraction(self, data_loader):
        """part_based_engine.py:132-166 with the features kept on the device.  `data_loader` yields dicts with
        'image', 'pid', 'camid' (and optionally 'mask'), like the reference's loaders."""
        f_, pids_, camids_, vis_, masks_ = [], [], [], [], []
        dev = next(self.model.parameters()).device
        for data in data_loader:
            imgs = data["image"].to(dev, non_blocking=True)
            ext = data.get("mask")
            if ext is not None:
                ext = ext.to(dev, non_blocking=True)
            out = self.model(imgs, external_parts_masks=ext)
            features, visibility_scores, parts_masks, _ = self.extract_test_embeddings(out)
            f_.append(features)
            masks_.append(parts_masks)
            if self.mask_filtering_testing:
                vis_.append(visibility_scores)
            pids_.extend(np.asarray(data["pid"]).tolist())
            camids_.extend(np.asarray(data["camid"]).tolist())
        f_ = torch.cat(f_, 0)
        vis_ = torch.cat(vis_, 0) if self.mask_filtering_testing else None
        return f_, np.asarray(pids_), np.asarray(camids_), vis_, torch.cat(masks_, 0)

    @torch.no_grad()
    def _evaluate(self, query_loader, gallery_loader, dist_metric="euclidean", normalize_feature=False, max_rank=50,
                  rerank=False, exact_distance=False):
        """Feature extraction -> part distance -> CMC/mAP, device-resident end to end; returns (cmc ndarray, mAP)."""
        qf, q_pids, q_camids, qv, _ = self._feature_extraction(query_loader)
        gf, g_pids, g_camids, gv, _ = self._feature_extraction(gallery_loader)
        if normalize_feature:
            qf, gf = self.normalize(qf), self.normalize(gf)
        impl = 1 if exact_distance else 0
        dist, part = metrics.part_distance_device(qf, gf, qv, gv, self.dist_combine_strat, dist_metric, impl=impl)
        if rerank:                                                   # part_based_engine.py:218-226
            from .rerank import re_ranking_device
            qq, _ = metrics.part_distance_device(qf, qf, qv, qv, self.dist_combine_strat, dist_metric, impl=impl,
                                                 want_part_dist=False)
            gg, _ = metrics.part_distance_device(gf, gf, gv, gv, self.dist_combine_strat, dist_metric, impl=impl,
                                                 want_part_dist=False)
            dist = re_ranking_device(dist, qq, gg)
        res = metrics.rank_device(dist, q_pids, g_pids, q_camids, g_camids, max_rank)
        metrics._check_rank(res, dist.shape[1], max_rank)
        self.last_distmat, self.last_body_parts_distmat = dist, part
        return res["cmc"].cpu().numpy(), float(res["mAP"].item())

    def individual_parts_ranking_performances(self, body_parts_distmat, q_pids, g_pids, q_camids, g_camids, max_rank=50):
        """display_individual_parts_ranking_performances (part_based_engine.py:308-339) without the printing: one
        device-resident CMC/mAP evaluation per test embedding; returns rows [title, mAP, R-1, R-5, R-10]."""
        emb = self.test_embeddings
        offset = int("globl" in emb) + int("foreg" in emb)
        table = []
        for bp in range(body_parts_distmat.shape[0]):
            res = metrics.rank_device(body_parts_distmat[bp], q_pids, g_pids, q_camids, g_camids, max_rank)
            metrics._check_rank(res, body_parts_distmat.shape[2], max_rank)
            cmc = res["cmc"].cpu().numpy()
            title = "p {}".format(bp - offset)
            if bp < offset:
                title = ("globl" if "globl" in emb else "foreg") if bp == 0 else "foreg"
            table.append([title, float(res["mAP"].item()), float(cmc[0]), float(cmc[min(4, len(cmc) - 1)]),
                          float(cmc[min(9, len(cmc) - 1)])])
        return table

    # ------------------------------------------------------------------ training
    def combine_losses(self, visibility_scores_dict, embeddings_dict, id_cls_scores_dict, pids, pixels_cls_scores=None,
                       target_masks=None, bpa_weight=0):
        """part_based_engine.py:107-130 (the bilinear resize + argmax of the target masks is fused into the kernel)."""
        loss, loss_summary = self.GiLt(embeddings_dict, visibility_scores_dict, id_cls_scores_dict, pids)
        if pixels_cls_scores is not None and target_masks is not None and bpa_weight > 0:
            bpa_loss, bpa_summary = self.body_part_attention_loss.forward_from_masks(pixels_cls_scores, target_masks)
            loss = loss + bpa_weight * bpa_loss
            loss_summary = {**loss_summary, **bpa_summary}
        return loss, loss_summary

    def forward_backward(self, data):
        """part_based_engine.py:77-105: forward, GiLt (+ body-part attention) loss, backward, optimizer step -- one
        fused device-side training step (bpbreid_b200.train.TrainStep)."""
        if getattr(self, "_train_step", None) is None:
            from .train import TrainStep
            bpa_w = 0.0
            if self.losses_weights is not None and PIXELS in self.losses_weights:
                bpa_w = float(self.losses_weights[PIXELS].get("ce", 0.0))
            self._train_step = TrainStep(self.model, self.GiLt, self.optimizer, bpa_weight=bpa_w)
        dev = next(self.model.parameters()).device
        imgs = data["image"].to(dev, non_blocking=True)
        pids = data["pid"].to(dev, non_blocking=True)
        masks = data.get("mask")
        if masks is not None:
            masks = masks.to(dev, non_blocking=True)
            assert masks.shape[1] == self.parts_num + 1
        return self._train_step(imgs, pids, masks)
