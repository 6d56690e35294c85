"""ORACLE tooling (test infrastructure only): mints the golden vectors under tests/golden/ by running the
UNMODIFIED reference modules (imported from /root/reference through oracle/ref_loader.py) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):   python oracle/make_golden.py
Everything it writes is small (< 1 MB per file); model weights are NOT stored -- they are regenerated from a seed
with bpbreid_b200.params.seeded_state_dict (deterministic CPU RNG, same torch build on both machines).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tests.golden_inputs import (distance_case, loss_case, model_case, rank_case, LOSS_CASES, DIST_CASES,  # noqa: E402
                                 RANK_CASES, MODEL_CASES)

OUT = os.path.join(ROOT, "tests", "golden")


def t2n(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def gen_losses(ref):
    out = {}
    for name in LOSS_CASES:
        c = loss_case(name)
        g = ref.GiLtLoss(c["weights"], use_visibility_scores=c["use_vis"], triplet_margin=c["margin"],
                         loss_name="part_averaged_triplet_loss", writer=ref_loader.NoOpWriter())
        emb = {k: torch.from_numpy(v) for k, v in c["emb"].items()}
        vis = {k: torch.from_numpy(v) for k, v in c["vis"].items()}
        ids = {k: torch.from_numpy(v) for k, v in c["ids"].items()}
        pids = torch.from_numpy(c["pids"])
        loss, summary = g(emb, vis, ids, pids)
        out[f"{name}/loss"] = t2n(loss)
        for key, info in summary.items():
            for kk, vv in info.items():
                out[f"{name}/{key}/{kk}"] = t2n(vv).astype(np.float64)
        # the triplet building blocks on their own
        tl = ref.losses.PartAveragedTripletLoss(margin=c["margin"], writer=ref_loader.NoOpWriter())
        pd = tl._part_based_pairwise_distance_matrix(emb["parts"].transpose(1, 0), squared=False)
        out[f"{name}/part_pairwise"] = t2n(pd)
    # body part attention loss
    torch.manual_seed(7)
    logits = torch.randn(3, 6, 16, 8)
    masks = torch.softmax(3 * torch.randn(3, 6, 9, 5), 1)
    tm = torch.nn.functional.interpolate(masks, logits.shape[2:], mode="bilinear", align_corners=True)
    tgt = tm.argmax(1)
    bpa = ref.BodyPartAttentionLoss(loss_type="cl", label_smoothing=0.1)
    l, s = bpa(logits, tgt)
    out["bpa/logits"], out["bpa/masks"] = t2n(logits), t2n(masks)
    out["bpa/loss"], out["bpa/acc"] = t2n(l), np.float64(s["pixls"]["a"])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)
    print("losses.npz", len(out))


def gen_distance(ref):
    out = {}
    for name in DIST_CASES:
        c = distance_case(name)
        qv = None if c["qv"] is None else torch.from_numpy(c["qv"])
        gv = None if c["gv"] is None else torch.from_numpy(c["gv"])
        d, pd = ref.distance.compute_distance_matrix_using_bp_features(
            torch.from_numpy(c["qf"]), torch.from_numpy(c["gf"]), qv, gv, c["strat"], 17, False, c["metric"])
        out[f"{name}/dist"], out[f"{name}/part_dist"] = t2n(d), t2n(pd)
    np.savez_compressed(os.path.join(OUT, "distance.npz"), **out)
    print("distance.npz", len(out))


def gen_rank(ref):
    out = {}
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    try:
        import rank_cy  # the reference's own Cython module, built by oracle/build.py --ref
    except ImportError:
        rank_cy = None
    for name in RANK_CASES:
        c = rank_case(name)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            r = ref.rank.eval_market1501(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"], c["max_rank"])
        out[f"{name}/cmc"], out[f"{name}/mAP"] = np.asarray(r["cmc"]), np.float64(r["mAP"])
        # same unmodified reference code with numpy's sort pinned to the stable kind (the ranking contract:
        # the reference's default introsort is free to permute tie groups, SURVEY.md section 7 hard part 5)
        orig = np.argsort
        np.argsort = lambda a, axis=-1, kind=None, order=None: orig(a, axis=axis, kind="stable")
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                rs_ = ref.rank.eval_market1501(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"],
                                               c["max_rank"])
            out[f"{name}/cmc_stable"], out[f"{name}/mAP_stable"] = np.asarray(rs_["cmc"]), np.float64(rs_["mAP"])
            if rank_cy is not None:
                with contextlib.redirect_stdout(io.StringIO()):
                    cmc_s, map_s = rank_cy.evaluate_cy(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"],
                                                       c["g_camids"], c["max_rank"])
                out[f"{name}/cy_cmc_stable"], out[f"{name}/cy_mAP_stable"] = np.asarray(cmc_s), np.float64(map_s)
        finally:
            np.argsort = orig
        if rank_cy is not None:
            with contextlib.redirect_stdout(io.StringIO()):
                cmc, m_ap = rank_cy.evaluate_cy(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"],
                                                c["max_rank"])
            out[f"{name}/cy_cmc"], out[f"{name}/cy_mAP"] = np.asarray(cmc), np.float64(m_ap)
    np.savez_compressed(os.path.join(OUT, "rank.npz"), **out)
    print("rank.npz", len(out))


def gen_rerank():
    """tests/golden/rerank.npz: the UNMODIFIED reference re-ranking (torchreid/utils/rerank.py, numpy only, imported by
    path from where it lies) on the seeded cases of tests/golden_inputs.py::rerank_case."""
    import importlib.util
    from tests.golden_inputs import RERANK_CASES, rerank_case
    spec = importlib.util.spec_from_file_location("ref_rerank", os.path.join(ref_loader.REF_PKG, "utils", "rerank.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for name in RERANK_CASES:
        c = rerank_case(name)
        out[f"{name}/final"] = mod.re_ranking(c["q_g"], c["q_q"], c["g_g"], k1=c["k1"], k2=c["k2"],
                                              lambda_value=c["lam"]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "rerank.npz"), **out)
    print("rerank.npz", len(out))


def gen_inputs():
    """tests/golden/inputs.npz: the reference's mask transforms (mask_transform.py / pifpaf_mask_transform.py, imported
    from where they lie with a stub for the absent albumentations base class) composed in build_transforms' order
    (transforms.py:73-163) with cv2.resize doing what albumentations.Resize does; images: cv2.resize + the restated
    albumentations.Normalize.  Also stores the reference's group tables so the product's restated tables are checked."""
    import importlib
    import types

    import cv2
    from oracle import input_oracle as IO
    from tests.golden_inputs import INPUT_CASES, input_case
    alb = types.ModuleType("albumentations")

    class DualTransform:                                        # the only albumentations name mask_transform.py uses
        def __init__(self, always_apply=False, p=0.5):
            pass
    alb.DualTransform = DualTransform
    sys.modules["albumentations"] = alb
    for name, sub in (("torchreid.data", "data"), ("torchreid.data.masks_transforms", "data/masks_transforms")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(ref_loader.REF_PKG, sub)]
        m.__package__ = name
        sys.modules[name] = m
    mt = importlib.import_module("torchreid.data.masks_transforms.mask_transform")
    pp = importlib.import_module("torchreid.data.masks_transforms.pifpaf_mask_transform")
    classes = {"five_v": pp.CombinePifPafIntoFiveVerticalParts, "eight": pp.CombinePifPafIntoEightBodyMasks}
    out = {}
    for prep, cls in classes.items():
        t = cls()
        tab = np.zeros((t.parts_num, 36), np.uint8)
        for i, part in enumerate(t.parts_names):
            for k in t.parts_grouping[part]:
                tab[i, t.parts_map[k]] = 1
        out[f"groups/{prep}"] = tab
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    for name in INPUT_CASES:
        c = input_case(name)
        h, w = c["height"], c["width"]
        group = classes[c["preprocess"]]()
        addbg = mt.AddBackgroundMask(c["strategy"], c["softmax_weight"], c["threshold"])
        resize = mt.ResizeMasks(h, w, 4)
        imgs, masks = [], []
        for img, f, p in zip(c["images"], c["fields"], c["params"]):
            x = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
            m = np.transpose(f, (1, 2, 0))                                  # read_masks: HWC (utils/tools.py:139-140)
            m = np.stack([cv2.resize(np.ascontiguousarray(m[:, :, k]), (w, h), interpolation=cv2.INTER_NEAREST)
                          for k in range(36)], axis=2)
            if p["flip"]:
                x, m = np.ascontiguousarray(x[:, ::-1]), np.ascontiguousarray(m[:, ::-1])
            if p["crop"] is not None:
                x = cv2.copyMakeBorder(x, 10, 10, 10, 10, cv2.BORDER_CONSTANT, value=0)
                m = np.pad(m, ((10, 10), (10, 10), (0, 0)))
                cx, cy = p["crop"]
                x, m = x[cy:cy + h, cx:cx + w], m[cy:cy + h, cx:cx + w]
            xf = IO.normalize(np.ascontiguousarray(x), mean, std)
            if p["erase"] is not None:
                x0, y0, x1, y1 = p["erase"]
                xf[y0:y1, x0:x1] = np.array([0.485, 0.456, 0.406], np.float32)
                m = m.copy()
                m[y0:y1, x0:x1] = 0
            imgs.append(xf.transpose(2, 0, 1))
            t = mt.PermuteMasksDim().apply_to_mask(torch.as_tensor(np.ascontiguousarray(m)))
            t = resize.apply_to_mask(addbg.apply_to_mask(group.apply_to_mask(t)))
            masks.append(t.numpy())
        out[f"{name}/images"], out[f"{name}/masks"] = np.stack(imgs).astype(np.float32), np.stack(masks).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "inputs.npz"), **out)
    print("inputs.npz", len(out))


def gen_models(ref, only=None):
    for name in (only or MODEL_CASES):
        c = model_case(name)
        m = ref_loader.build_reference_model(c["backbone"], c["K"], c["num_classes"])
        missing = m.load_state_dict(c["state_dict"], strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        out = {}
        for mode in ("train", "eval"):
            m.train(mode == "train")
            # fresh copy of running stats each time so that train-mode side effects do not leak into eval
            m.load_state_dict(c["state_dict"], strict=True)
            with torch.no_grad():
                emb, vis, ids, pix, feats, masks = m(torch.from_numpy(c["images"]))
            for k, v in emb.items():
                out[f"{mode}/emb/{k}"] = t2n(v)
            for k, v in vis.items():
                out[f"{mode}/vis/{k}"] = t2n(v)
            for k, v in ids.items():
                out[f"{mode}/ids/{k}"] = t2n(v)
            out[f"{mode}/pix"] = t2n(pix)
            out[f"{mode}/feat_mean"] = t2n(feats.mean(dim=(2, 3)))
            out[f"{mode}/feat_abs_mean"] = t2n(feats.abs().mean())
            if mode == "train":
                g = ref.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3,
                                 loss_name="part_averaged_triplet_loss", writer=ref_loader.NoOpWriter())
                loss, _ = g(emb, vis, ids, torch.from_numpy(c["pids"]))
                out["train/gilt_loss"] = t2n(loss)
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)
        print(f"model_{name}.npz", len(out))


def gen_mask_models(ref):
    """eval-mode outputs of the reference with external part masks (bpbreid.py:149-175)."""
    from tests.golden_inputs import MASK_CASES, mask_case
    for name in MASK_CASES:
        c = mask_case(name)
        m = ref_loader.build_reference_model(c["backbone"], c["K"], c["num_classes"], **c["over"])
        m.load_state_dict(c["state_dict"], strict=True)
        m.eval()
        with torch.no_grad():
            emb, vis, ids, pix, feats, masks = m(torch.from_numpy(c["images"]),
                                                 external_parts_masks=torch.from_numpy(c["masks"]))
        out = {}
        for k, v in emb.items():
            out[f"eval/emb/{k}"] = t2n(v)
        for k, v in vis.items():
            out[f"eval/vis/{k}"] = t2n(v)
        for k, v in masks.items():
            out[f"eval/masks/{k}"] = t2n(v).astype(np.float32)
        out["eval/has_pix"] = np.int64(pix is not None)
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)
        print(f"model_{name}.npz", len(out))


GRAD_FULL_KEYS = ("pixel_classifier.classifier.weight", "pixel_classifier.classifier.bias", "pixel_classifier.bn.weight",
                  "pixel_classifier.bn.bias", "global_identity_classifier.bn.weight",
                  "backbone_appearance_feature_extractor.conv1.weight", "backbone_appearance_feature_extractor.bn1.weight",
                  "backbone_appearance_feature_extractor.stage4.2.branches.0.3.conv2.weight",
                  "backbone_appearance_feature_extractor.stage2.0.fuse_layers.1.0.0.0.weight",
                  "backbone_appearance_feature_extractor.stage3.1.fuse_layers.0.2.0.weight",
                  "backbone_appearance_feature_extractor.incre_modules.3.0.conv3.weight",
                  "backbone_appearance_feature_extractor.layer1.0.downsample.0.weight",
                  "parts_after_pooling_dim_reduce.layers.0.bias", "parts_after_pooling_dim_reduce.layers.1.weight")


def gen_grads(ref, name="hrnet32_k5_ctr", bpa_weight=0.35):
    """Reference `loss.backward()` (part_based_engine.py:98-103) on the contractive HRNet case: GiLt (default weights)
    + 0.35 * body-part attention loss.  Stores every parameter's gradient L2 norm and a few full tensors."""
    c = model_case(name)
    m = ref_loader.build_reference_model(c["backbone"], c["K"], c["num_classes"])
    m.load_state_dict(c["state_dict"], strict=True)
    m.train()
    masks = torch.from_numpy(c["masks"])
    emb, vis, ids, pix, feats, pm = m(torch.from_numpy(c["images"]), external_parts_masks=masks)
    g = ref.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, loss_name="part_averaged_triplet_loss",
                     writer=ref_loader.NoOpWriter())
    pids = torch.from_numpy(c["pids"])
    loss, _ = g(emb, vis, ids, pids)
    tm = torch.nn.functional.interpolate(masks, pix.shape[2:], mode="bilinear", align_corners=True)
    bpa = ref.BodyPartAttentionLoss(loss_type="cl", label_smoothing=0.1)
    l2, _ = bpa(pix, tm.argmax(1))
    total = loss + bpa_weight * l2
    total.backward()
    out = {"loss": t2n(total), "gilt": t2n(loss), "bpa": t2n(l2)}
    for k, p in m.named_parameters():
        if p.grad is not None:
            out[f"norm/{k}"] = np.float64(p.grad.double().norm().item())
            if k in GRAD_FULL_KEYS:
                out[f"full/{k}"] = t2n(p.grad)
    np.savez_compressed(os.path.join(OUT, f"grads_{name}.npz"), **out)
    print(f"grads_{name}.npz", len(out), float(total))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    which = sys.argv[1:] or ["losses", "distance", "rank", "rerank", "inputs", "models", "grads", "masks"]
    if "losses" in which:
        gen_losses(ref)
    if "distance" in which:
        gen_distance(ref)
    if "rank" in which:
        gen_rank(ref)
    if "rerank" in which:
        gen_rerank()
    if "inputs" in which:
        gen_inputs()
    if "models" in which:
        gen_models(ref)
    only = [w.split(":", 1)[1] for w in which if w.startswith("models:")]
    if only:
        gen_models(ref, only)
    if "grads" in which:
        gen_grads(ref)
    if "masks" in which:
        gen_mask_models(ref)
