"""ORACLE tooling (test infrastructure only).

Imports the UNMODIFIED reference hot-path modules from /root/reference so that (a) oracle/make_golden.py can
mint golden vectors from the reference itself and (b) tests can cross-check the restatements in this
directory against it *in this container*.  /root/reference does not exist on the GPU box: everything that
runs there uses the committed fixtures under tests/golden/ instead.  Nothing is copied: the reference files
are imported from where they lie.

Recipe: `import torchreid` fails (albumentations / matplotlib / yacs absent), so the
package objects `torchreid`, `torchreid.utils`, `torchreid.metrics` are registered as bare namespace stubs
whose __path__ points into the reference tree -- their __init__.py is never executed -- and the leaf
modules (models, losses, metrics.distance, metrics.rank, utils.tensortools, utils.constants) import by path.
Third-party names that only feed logging are stubbed (yacs CfgNode, torchmetrics.Accuracy, monai losses,
the Writer singleton).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("BPB_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "torchreid")


def available() -> bool:
    return os.path.isdir(REF_PKG)


class CfgNode(dict):
    """Minimal stand-in for yacs.config.CfgNode: attribute access over a dict (hrnet.py:11,26-56)."""

    def __init__(self, init=None, new_allowed=False):
        super().__init__(init or {})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class NoOpWriter:
    """The three Writer hooks the losses call (part_averaged_triplet_loss.py:69,138)."""

    def update_invalid_pairwise_distances_count(self, *a, **k):
        pass

    def update_invalid_part_based_pairwise_distances_count(self, *a, **k):
        pass

    def used_parts_statistics(self, *a, **k):
        pass


_loaded = None


def _stub_pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with .models, .losses, .distance, .rank, .tensortools (reference modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    yacs = types.ModuleType("yacs")
    yacs_cfg = types.ModuleType("yacs.config")
    yacs_cfg.CfgNode = CfgNode
    sys.modules.setdefault("yacs", yacs)
    sys.modules.setdefault("yacs.config", yacs_cfg)

    class Accuracy(torch.nn.Module):           # torchmetrics.Accuracy(top_k=1): logging only (GiLt_loss.py:37,118)
        def __init__(self, top_k=1):
            super().__init__()

        def forward(self, scores, target):
            return (scores.argmax(1) == target).float().mean()

    tm = types.ModuleType("torchmetrics")
    tm.Accuracy = Accuracy
    sys.modules.setdefault("torchmetrics", tm)
    monai = types.ModuleType("monai")
    monai_losses = types.ModuleType("monai.losses")
    monai_losses.FocalLoss = monai_losses.DiceLoss = object
    sys.modules.setdefault("monai", monai)
    sys.modules.setdefault("monai.losses", monai_losses)

    tr = _stub_pkg("torchreid", REF_PKG)
    tr.utils = _stub_pkg("torchreid.utils", os.path.join(REF_PKG, "utils"))
    wr = types.ModuleType("torchreid.utils.writer")
    wr.Writer = type("Writer", (), {"current_writer": classmethod(lambda cls: None)})
    sys.modules["torchreid.utils.writer"] = wr
    tr.utils.writer = wr

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        tr.models = importlib.import_module("torchreid.models")
        tr.losses = importlib.import_module("torchreid.losses")
    tr.metrics = _stub_pkg("torchreid.metrics", os.path.join(REF_PKG, "metrics"))
    distance = importlib.import_module("torchreid.metrics.distance")
    spec = importlib.util.spec_from_file_location("ref_rank", os.path.join(REF_PKG, "metrics", "rank.py"))
    rank = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(rank)
    tensortools = importlib.import_module("torchreid.utils.tensortools")
    gilt = importlib.import_module("torchreid.losses.GiLt_loss")
    bpa = importlib.import_module("torchreid.losses.body_part_attention_loss")

    ns = types.SimpleNamespace(models=tr.models, losses=tr.losses, distance=distance, rank=rank,
                               tensortools=tensortools, GiLtLoss=gilt.GiLtLoss,
                               BodyPartAttentionLoss=bpa.BodyPartAttentionLoss)
    _loaded = ns
    return ns


def make_cfg(backbone="hrnet32", parts_num=5, **over):
    """cfg.model.bpbreid node needed by bpbreid(...) (bpbreid.py:22-47); defaults = default_config.py:45-58."""
    b = CfgNode(dict(pooling="gwap", normalization="identity", mask_filtering_training=False,
                     mask_filtering_testing=True, last_stride=1, dim_reduce="after_pooling",
                     dim_reduce_output=512, backbone=backbone, learnable_attention_enabled=True,
                     test_embeddings=["bn_foreg", "parts"], test_use_target_segmentation="none",
                     training_binary_visibility_score=True, testing_binary_visibility_score=True,
                     shared_parts_id_classifier=False, hrnet_pretrained_path="",
                     masks=CfgNode(dict(parts_num=parts_num))))
    b.update(over)
    return CfgNode(dict(model=CfgNode(dict(bpbreid=b))))


def build_reference_model(backbone="hrnet32", parts_num=5, num_classes=751, **over):
    ref = load()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return ref.models.build_model("bpbreid", num_classes=num_classes, loss="part_based", pretrained=False,
                                      use_gpu=False, config=make_cfg(backbone, parts_num, **over))
