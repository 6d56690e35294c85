"""Host-side mirror of the reference model-builder API for the BPBReID hot path.

    build_model(name, num_classes, loss='part_based', pretrained, use_gpu, config=cfg)   torchreid/models/__init__.py:102-134
    bpbreid(num_classes, loss, pretrained, config)                                        torchreid/models/bpbreid.py:510-518
    BPBreID.forward(images, external_parts_masks) -> 6-tuple of dicts                     torchreid/models/bpbreid.py:116-259

Same names, argument meaning and error behaviour (KeyError for unknown model names, ValueError for unknown
pooling / normalisation, swallowed extra kwargs); ``state_dict()`` keys are identical to the reference's so its
checkpoints load unchanged.  The arithmetic runs in libbpbreid_sm100.so through `bpbreid_b200.plan.ForwardPlan`;
there is NO CPU / eager-PyTorch fallback: calling the model with CPU tensors raises.
"""
from __future__ import annotations

import torch
from torch import nn

from . import params as P
from .constants import *  # noqa: F401,F403  (GLOBAL, FOREGROUND, ... same strings as torchreid/utils/constants.py)

__all__ = ["build_model", "bpbreid", "BPBreID", "show_avai_models"]


class BPBreID(nn.Module):
    """Body-part-based re-id network (parameter container + CUDA forward plan)."""

    def __init__(self, num_classes, pretrained, loss, model_cfg, horizontal_stripes=False, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_classes = num_classes
        self.parts_num = model_cfg.masks.parts_num
        self.horizontal_stripes = horizontal_stripes
        self.shared_parts_id_classifier = model_cfg.shared_parts_id_classifier
        self.test_use_target_segmentation = model_cfg.test_use_target_segmentation
        self.training_binary_visibility_score = model_cfg.training_binary_visibility_score
        self.testing_binary_visibility_score = model_cfg.testing_binary_visibility_score
        self.learnable_attention_enabled = model_cfg.learnable_attention_enabled
        if model_cfg.pooling not in ("gap", "gmp", "gwap"):
            raise ValueError("pooling type {} not supported".format(model_cfg.pooling))
        if model_cfg.normalization not in ("identity", "batch_norm_3d", "batch_norm_2d", "batch_norm_1d"):
            raise ValueError("normalization type {} not supported".format(model_cfg.normalization))
        if model_cfg.backbone not in P.BACKBONES:
            raise KeyError("Unknown model: {}. Must be one of {}".format(model_cfg.backbone, list(P.BACKBONES)))
        if pretrained:
            # the reference downloads / reads ImageNet weights here (hrnet.py:588-617); there is no network and the
            # checkpoint files are not shipped: load them explicitly with load_state_dict instead.
            raise FileNotFoundError("pretrained backbone weights are not available offline; build with "
                                    "pretrained=False and load a state_dict")
        self.backbone_appearance_feature_extractor = P.BACKBONES[model_cfg.backbone](
            num_classes, dim_reduction_channels=model_cfg.dim_reduce_output)
        c = self.backbone_appearance_feature_extractor.feature_dim
        d = model_cfg.dim_reduce_output
        self.spatial_feature_size = c
        self.dim_reduce_output = d
        self.after_pooling_dim_reduce = model_cfg.dim_reduce == "after_pooling"
        if not self.after_pooling_dim_reduce:
            raise NotImplementedError("dim_reduce='{}': only 'after_pooling' (every shipped config) is built"
                                      .format(model_cfg.dim_reduce))
        self.global_after_pooling_dim_reduce = P.DimReduceParams(c, d)
        self.foreground_after_pooling_dim_reduce = P.DimReduceParams(c, d)
        self.background_after_pooling_dim_reduce = P.DimReduceParams(c, d)
        self.parts_after_pooling_dim_reduce = P.DimReduceParams(c, d)
        self.pixel_classifier = P.PixelClassifierParams(c, self.parts_num)
        self.global_identity_classifier = P.BNClassifierParams(d, num_classes)
        self.background_identity_classifier = P.BNClassifierParams(d, num_classes)
        self.foreground_identity_classifier = P.BNClassifierParams(d, num_classes)
        self.concat_parts_identity_classifier = P.BNClassifierParams(self.parts_num * d, num_classes)
        if self.shared_parts_id_classifier:
            self.parts_identity_classifier = P.BNClassifierParams(d, num_classes)
        else:
            self.parts_identity_classifier = nn.ModuleList(
                [P.BNClassifierParams(d, num_classes) for _ in range(self.parts_num)])
        self._plans = {}
        self.conv_impl = "tc"
        self.use_cuda_graph = False
        self.debug_trace = False
        # HRNet only: evaluate the head on the un-concatenated branch maps (exact linear-algebra identity, see
        # csrc/head_lowres.cu).  `materialize_spatial_features=False` additionally skips writing the [N,1920,H/4,W/4]
        # map that is then only an *output* (5th element of the forward tuple becomes None).
        self.lowres_head = True
        self.materialize_spatial_features = True
        # inference: BatchNorm / residual / ReLU ride the convolution's epilogue (bit-identical to conv + bn_apply)
        self.fuse_eval_epilogue = True

    # -- forward ---------------------------------------------------------------------------------------------
    def _needs_masks(self):
        return (not self.learnable_attention_enabled) or \
               (not self.training and self.test_use_target_segmentation in ("hard", "soft"))

    def plan_for(self, images, external_parts_masks=None):
        if not images.is_cuda:
            raise RuntimeError("bpbreid_b200.BPBreID runs on sm_100a only: move the model and the images to a CUDA "
                               "device (there is no CPU fallback path)")
        n, _, h, w = images.shape
        ext_shape = tuple(external_parts_masks.shape) if (self._needs_masks() and external_parts_masks is not None) else None
        key = (n, h, w, self.training, images.device.index, self.conv_impl, self.use_cuda_graph, self.debug_trace,
               self.lowres_head, self.materialize_spatial_features, ext_shape, self.fuse_eval_epilogue)
        plan = self._plans.get(key)
        if plan is None:
            from .plan import ForwardPlan
            p0 = next(self.parameters())
            if p0.device != images.device:
                raise RuntimeError(f"model parameters live on {p0.device}, images on {images.device}")
            plan = ForwardPlan(self, n, h, w, self.training, images.device, self.conv_impl, self.use_cuda_graph,
                               self.debug_trace, ext_shape=ext_shape)
            self._plans[key] = plan
        return plan

    def forward(self, images, external_parts_masks=None):
        if not self.learnable_attention_enabled:
            assert external_parts_masks is not None                       # bpbreid.py:151
        if not self.training and self.test_use_target_segmentation in ("hard", "soft"):
            assert external_parts_masks is not None                       # bpbreid.py:162,171
        plan = self.plan_for(images, external_parts_masks)
        if plan.ext_mode:
            plan.run(images, external_parts_masks.to(device=images.device, dtype=torch.float32))
        else:
            plan.run(images)
        return plan.outputs()

    def invalidate_plans(self):
        """Call after the weights changed (optimizer step / load_state_dict): bf16 weight copies are re-packed."""
        for p in self._plans.values():
            p.refresh_weights()

    def _apply(self, fn, *a, **k):
        # .cuda() / .to() re-allocates parameters: cached plans hold raw device pointers and must be dropped
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_plans()
        self._param_version = getattr(self, "_param_version", 0) + 1      # training plans re-pack their operand copies
        return out


def bpbreid(num_classes, loss="part_based", pretrained=True, config=None, **kwargs):
    return BPBreID(num_classes, pretrained, loss, config.model.bpbreid, **kwargs)


__model_factory = {"bpbreid": bpbreid}


def show_avai_models():
    print(list(__model_factory.keys()))


def build_model(name, num_classes, loss="softmax", pretrained=True, use_gpu=True, **kwargs):
    avai_models = list(__model_factory.keys())
    if name not in avai_models:
        raise KeyError("Unknown model: {}. Must be one of {}".format(name, avai_models))
    return __model_factory[name](num_classes=num_classes, loss=loss, pretrained=pretrained, use_gpu=use_gpu, **kwargs)
