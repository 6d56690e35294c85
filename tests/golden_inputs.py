"""Seeded input generators shared by oracle/make_golden.py (which runs the reference on them, in the build
container) and the parity tests (which run the oracle restatements and the CUDA path on the same inputs).
Only numpy / torch CPU RNGs: bit-identical on the build container and on the GPU box (same image)."""
from __future__ import annotations

import numpy as np
import torch

F32 = np.float32
GL, FG, CC, PA = "globl", "foreg", "conct", "parts"

DEFAULT_W = {GL: {"id": 1., "tr": 0.}, FG: {"id": 1., "tr": 0.}, CC: {"id": 1., "tr": 0.}, PA: {"id": 0., "tr": 1.}}
MARKET_W = {GL: {"id": 1., "tr": 0.}, FG: {"id": 1., "tr": 1.}, CC: {"id": 1., "tr": 0.}, PA: {"id": 0., "tr": 1.}}

LOSS_CASES = ["default", "market", "vis_bool", "vis_float", "dups_singletons", "soft_margin"]


def loss_case(name, n=32, k=5, d=64, ncls=41):
    rs = np.random.RandomState({"default": 1, "market": 2, "vis_bool": 3, "vis_float": 4, "dups_singletons": 5,
                                "soft_margin": 6}[name])
    pids = (np.arange(n) // 4).astype(np.int64)
    parts = rs.randn(n, k, d).astype(F32)
    c = dict(weights=DEFAULT_W, use_vis=False, margin=0.3)
    if name == "market":
        c["weights"] = MARKET_W
    if name == "soft_margin":
        c["margin"] = 0.0
    vis_parts = np.ones((n, k), dtype=bool)
    vis_h = np.ones(n, dtype=bool)
    if name == "vis_bool":
        c["use_vis"] = True
        c["weights"] = MARKET_W
        vis_parts = rs.rand(n, k) > 0.35
        vis_parts[3] = False                     # a sample with no visible part at all -> -1 sentinels
        vis_parts[8, :] = [True, False, False, False, False]
        vis_parts[9, :] = [False, True, False, False, False]   # 8 and 9 share an id but no visible part
        vis_h = vis_parts.any(1)
    if name == "vis_float":
        c["use_vis"] = True
        vis_parts = rs.rand(n, k).astype(F32)
        vis_parts[5] = 0
        vis_h = vis_parts.max(1)
    if name == "dups_singletons":
        parts[1] = parts[0]                      # exact duplicates -> exact-zero distance mask
        parts[6] = parts[2]                      # duplicate across identities
        pids[-4:] = [100, 101, 102, 103]         # singleton identities -> anchors without positives
    emb = {GL: rs.randn(n, d).astype(F32), FG: rs.randn(n, d).astype(F32), PA: parts,
           CC: parts.reshape(n, k * d).copy()}
    ids = {GL: rs.randn(n, ncls).astype(F32), FG: rs.randn(n, ncls).astype(F32), CC: rs.randn(n, ncls).astype(F32),
           PA: rs.randn(n, k, ncls).astype(F32)}
    vis = {GL: np.ones(n, dtype=vis_h.dtype), FG: vis_h, CC: vis_h, PA: vis_parts}
    pid_cls = (pids % ncls).astype(np.int64)
    c.update(emb=emb, ids=ids, vis=vis, pids=pid_cls)
    return c


DIST_CASES = ["novis_mean", "novis_max", "bool_mean", "bool_max", "float_mean", "bool_mean_cosine", "ragged"]


def distance_case(name):
    seed = DIST_CASES.index(name) + 11
    rs = np.random.RandomState(seed)
    nq, ng, p, d = (37, 91, 6, 64) if name != "ragged" else (5, 131, 3, 96)
    qf = rs.randn(nq, p, d).astype(F32)
    gf = rs.randn(ng, p, d).astype(F32)
    qf /= np.linalg.norm(qf, axis=-1, keepdims=True)
    gf /= np.linalg.norm(gf, axis=-1, keepdims=True)
    gf[7] = qf[3]                                  # exact duplicates -> zero distances / ties
    gf[8] = qf[3]
    c = dict(qf=qf, gf=gf, qv=None, gv=None, strat="mean", metric="euclidean")
    if "max" in name:
        c["strat"] = "max"
    if "cosine" in name:
        c["metric"] = "cosine"
    if name.startswith("bool") or name == "ragged":
        qv, gv = rs.rand(nq, p) > 0.3, rs.rand(ng, p) > 0.3
        qv[1] = False                              # query with nothing visible -> whole row is the sentinel
        gv[4] = False
        qv[2] = [True] + [False] * (p - 1)
        gv[5] = [False] + [True] * (p - 1)         # (2,5): no common visible part
        c["qv"], c["gv"] = qv, gv
    if name.startswith("float"):
        qv, gv = rs.rand(nq, p).astype(F32), rs.rand(ng, p).astype(F32)
        qv[1] = 0
        c["qv"], c["gv"] = qv, gv
    return c


RANK_CASES = ["random", "ties", "small_gallery", "some_invalid"]


def rank_case(name):
    rs = np.random.RandomState(RANK_CASES.index(name) + 31)
    nq, ng, max_rank = 60, 400, 50
    if name == "small_gallery":
        nq, ng = 12, 30                           # max_rank > Ng
    dist = rs.rand(nq, ng).astype(F32)
    npid = 25
    q_pids, g_pids = rs.randint(0, npid, nq), rs.randint(0, npid, ng)
    q_cam, g_cam = rs.randint(0, 4, nq), rs.randint(0, 4, ng)
    if name == "small_gallery":
        q_cam[:] = 10                               # no same-camera removals: the reference's numpy path cannot
                                                    # stack rows shorter than max_rank (rank.py:152 would raise)
    if name == "ties":
        dist = np.round(dist * 8).astype(F32) / 8   # massive ties: only 9 distinct keys
        dist[:, :50] = 2.0                          # the "-1 -> max+1" plateau of distance.py:173-176
    if name == "some_invalid":
        q_pids[:5] = 1000                           # identities absent from the gallery
        sel = g_pids == q_pids[7]
        g_cam[sel] = q_cam[7]                       # only same-camera matches -> query skipped (rank.py:131-133)
    return dict(dist=dist, q_pids=q_pids.astype(np.int64), g_pids=g_pids.astype(np.int64),
                q_camids=q_cam.astype(np.int64), g_camids=g_cam.astype(np.int64), max_rank=max_rank)


# k-reciprocal re-ranking (utils/rerank.py:31-113): name -> (Nq, Ng, identities, k1, k2, lambda)
RERANK_CASES = {"clusters": (40, 160, 20, 20, 6, 0.3),        # the engine's defaults (part_based_engine.py:226)
                "small_k": (24, 100, 12, 5, 1, 0.5),           # k1/2 = 2.5 rounds to even, k2 == 1 skips the expansion
                "tiny": (6, 14, 4, 20, 6, 0.3)}                # k1 + 1 > Nq + Ng: every neighbour list is truncated


def rerank_case(name):
    """Euclidean distance matrices between L2-normalised clustered features (float32, no exact ties off the diagonal)."""
    nq, ng, nid, k1, k2, lam = RERANK_CASES[name]
    rs = np.random.RandomState(list(RERANK_CASES).index(name) + 71)
    centres = rs.randn(nid, 48).astype(F32)

    def feats(n):
        ids = rs.randint(0, nid, n)
        f = centres[ids] + 0.6 * rs.randn(n, 48).astype(F32)
        return (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(F32), ids

    def dist(a, b):
        d2 = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * (a @ b.T)
        return np.sqrt(np.maximum(d2, 0)).astype(F32)

    (qf, qid), (gf, gid) = feats(nq), feats(ng)
    return dict(q_g=dist(qf, gf), q_q=dist(qf, qf), g_g=dist(gf, gf), k1=k1, k2=k2, lam=lam, q_pids=qid, g_pids=gid)


# input pipeline (data/transforms.py:24-163, data/masks_transforms/mask_transform.py:21-85):
# name -> (H, W, masks preprocess, background strategy, threshold, softmax weight, augment?)
INPUT_CASES = {"test_384_five_v": (384, 128, "five_v", "threshold", 0.5, 15, False),
               "train_256_eight": (256, 128, "eight", "threshold", 0.5, 15, True),
               "test_256_diffmax": (256, 128, "five_v", "diff_from_max", 0.3, 15, False)}
_INPUT_SIZES = [((128, 64), (16, 8)), ((200, 77), (25, 10)), ((51, 23), (33, 17)), ((384, 128), (48, 16)),
                ((300, 150), (38, 19)), ((64, 32), (64, 32))]


def input_case(name):
    """A batch of RGB uint8 images of different sizes with their 36-channel pifpaf confidence stacks (float32, sparse)."""
    h, w, prep, strat, thr, smw, aug = INPUT_CASES[name]
    rs = np.random.RandomState(list(INPUT_CASES).index(name) + 91)
    images, fields, params = [], [], []
    for (ih, iw), (mh, mw) in _INPUT_SIZES:
        images.append(rs.randint(0, 256, (ih, iw, 3)).astype(np.uint8))
        f = rs.rand(36, mh, mw).astype(F32) * (rs.rand(36, mh, mw) > 0.6)
        fields.append(f.astype(F32))
        if aug:
            flip = bool(rs.rand() < 0.5)
            crop = (int(rs.randint(0, 21)), int(rs.randint(0, 21)))
            erase = None
            if rs.rand() < 0.7:                               # CoarseDropout geometry (transforms.py:112-117)
                eh, ew = rs.randint(int(h * 0.15), int(h * 0.65) + 1), rs.randint(int(w * 0.15), int(w * 0.65) + 1)
                y0, x0 = rs.randint(0, h - eh + 1), rs.randint(0, w - ew + 1)
                erase = (int(x0), int(y0), int(x0 + ew), int(y0 + eh))
            params.append(dict(flip=flip, crop=crop, erase=erase))
        else:
            params.append(dict(flip=False, crop=None, erase=None))
    return dict(images=images, fields=fields, params=params, height=h, width=w, preprocess=prep, strategy=strat,
                threshold=thr, softmax_weight=smw)


MODEL_CASES = ["hrnet32_k5", "resnet50_k5", "ibn_k8", "hrnet32_k5_384", "hrnet32_k5_ctr"]
# external part masks (bpbreid.py:149-175): name -> (base case, config overrides); eval-mode goldens with the masks passed
MASK_CASES = {"hrnet32_k5_extattn": ("hrnet32_k5", dict(learnable_attention_enabled=False)),
              "hrnet32_k5_hard": ("hrnet32_k5", dict(test_use_target_segmentation="hard")),
              "hrnet32_k5_soft": ("hrnet32_k5", dict(test_use_target_segmentation="soft")),
              "resnet50_k5_extattn": ("resnet50_k5", dict(learnable_attention_enabled=False)),
              "resnet50_k5_soft": ("resnet50_k5", dict(test_use_target_segmentation="soft"))}
# name -> (backbone, K, seed, H, W, contractive)
_MODEL_SPECS = {"hrnet32_k5": ("hrnet32", 5, 100, 256, 128, False), "resnet50_k5": ("resnet50", 5, 200, 256, 128, False),
                "ibn_k8": ("fastreid_resnet_ibn", 8, 300, 256, 128, False),
                # the reference's shipped input size (configs/bpbreid/*.yaml:5-6, BASELINE config 2)
                "hrnet32_k5_384": ("hrnet32", 5, 400, 384, 128, False),
                # "contractive" weights: the BatchNorm that closes every residual branch has gamma ~ 0.15, so each block is
                # a small perturbation of the identity and bf16 rounding noise is NOT amplified block after block; this
                # makes end-to-end TRAIN-mode comparisons (batch statistics) and end-to-end gradients well conditioned
                "hrnet32_k5_ctr": ("hrnet32", 5, 500, 256, 128, True)}


def contractive(sd):
    out = dict(sd)
    for k, v in sd.items():
        if k.endswith("bn2.weight") and ".branches." in k:
            out[k] = v * 0.15
        elif k.endswith("bn3.weight"):
            out[k] = v * 0.15
    return out


from bpbreid_b200.config import Cfg as _Cfg, make_cfg  # noqa: E402,F401  (one definition of the reference defaults)


def build_product_model(backbone, K, num_classes=751, **over):
    from bpbreid_b200 import models
    torch.manual_seed(1234)
    return models.build_model("bpbreid", num_classes=num_classes, loss="part_based", pretrained=False, use_gpu=False,
                              config=make_cfg(backbone, K, **over))


def model_case(name, n=None, num_classes=751, **over):
    """n defaults to 4, and to 16 for the contractive case: its purpose is end-to-end TRAIN-mode comparisons, where the
    head's BatchNorm1d layers normalise over the batch rows and need more than a handful of samples to be conditioned."""
    from bpbreid_b200 import params
    backbone, K, seed, h, w, ctr = _MODEL_SPECS[name]
    if n is None:
        n = 16 if ctr else 4
    m = build_product_model(backbone, K, num_classes, **over)
    sd = params.seeded_state_dict(m, seed)
    if ctr:
        sd = contractive(sd)
    g = torch.Generator().manual_seed(seed + 1)
    images = torch.randn(n, 3, h, w, generator=g)
    if ctr:
        # images that DIFFER from each other like photographs do (contrast, colour cast, large-scale structure).  Pure
        # white noise gives every image the same pooled statistics up to ~1/sqrt(H*W); the train-mode BatchNorm1d layers
        # of the head then normalise a between-image variation that is smaller than bf16 rounding noise, and any
        # end-to-end comparison measures that amplification instead of the kernels.
        scale = 0.4 + 1.6 * torch.rand(n, 1, 1, 1, generator=g)
        offs = torch.randn(n, 3, 1, 1, generator=g)
        fy, fx = 0.5 + 2.5 * torch.rand(n, 3, 1, 1, generator=g), 0.5 + 2.5 * torch.rand(n, 3, 1, 1, generator=g)
        ph = 6.2831853 * torch.rand(n, 3, 1, 1, generator=g)
        yy = torch.linspace(0, 1, h).view(1, 1, h, 1)
        xx = torch.linspace(0, 1, w).view(1, 1, 1, w)
        images = images * scale + offs + 1.5 * torch.sin(6.2831853 * (fy * yy + fx * xx) + ph)
    images = images.numpy()
    pids = (np.arange(n) // 2).astype(np.int64)
    # external part masks [N, K+1, H/8, W/8] (rows sum to 1, like AddBackgroundMask's softmax, mask_transform.py:81-84)
    masks = torch.softmax(3 * torch.randn(n, K + 1, h // 8, w // 8, generator=g), 1).numpy()
    return dict(backbone=backbone, K=K, num_classes=num_classes, state_dict=sd, images=images, pids=pids, model=m,
                masks=masks, over=over)


def mask_case(name):
    base, over = MASK_CASES[name]
    return model_case(base, **over)
