"""CPU: the oracle restatements (oracle/*.py, oracle/rank_oracle.c) against the golden vectors minted from the
UNMODIFIED reference (oracle/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO
from oracle import metrics_oracle as MO
from oracle import model_oracle as MDL
from tests import golden_inputs as GI


@pytest.mark.parametrize("name", GI.LOSS_CASES)
def test_gilt_oracle_matches_reference(golden, name):
    g = golden("losses")
    c = GI.loss_case(name)
    loss, summary = LO.gilt(c["emb"], c["vis"], c["ids"], c["pids"], c["weights"], c["use_vis"], c["margin"])
    np.testing.assert_allclose(loss, g[f"{name}/loss"], rtol=2e-5, atol=1e-6)
    for key, info in summary.items():
        for kk, vv in info.items():
            np.testing.assert_allclose(vv, g[f"{name}/{key}/{kk}"], rtol=2e-5, atol=1e-6, err_msg=f"{key}/{kk}")
    pd = LO.part_pairwise_dist(c["emb"]["parts"].transpose(1, 0, 2))
    ref = g[f"{name}/part_pairwise"]
    np.testing.assert_allclose(pd, ref, rtol=1e-4, atol=2e-3)      # sqrt amplifies 1-ulp differences near 0
    assert ((pd == 0) == (ref == 0)).mean() > 0.999


def test_body_part_attention_oracle(golden):
    g = golden("losses")
    loss, acc = LO.body_part_attention(g["bpa/logits"], g["bpa/masks"])
    np.testing.assert_allclose(loss, g["bpa/loss"], rtol=1e-5)
    np.testing.assert_allclose(acc, g["bpa/acc"], rtol=1e-6)


@pytest.mark.parametrize("name", GI.DIST_CASES)
def test_distance_oracle_matches_reference(golden, name):
    g = golden("distance")
    c = GI.distance_case(name)
    d, pd = MO.part_distance(c["qf"], c["gf"], c["qv"], c["gv"], c["strat"], c["metric"])
    np.testing.assert_allclose(d, g[f"{name}/dist"], rtol=1e-4, atol=2e-3)   # sqrt of a ~1-ulp d^2 at duplicates
    np.testing.assert_allclose(pd, g[f"{name}/part_dist"], rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("name", GI.RANK_CASES)
def test_rank_oracles_match_reference(golden, name):
    g = golden("rank")
    c = GI.rank_case(name)
    r = MO.eval_market1501(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"], c["max_rank"])
    # reference numpy path (fp64) with the sort pinned to stable: exact agreement, ties included
    np.testing.assert_array_equal(r["cmc"], g[f"{name}/cmc_stable"])
    np.testing.assert_allclose(r["mAP"], g[f"{name}/mAP_stable"], rtol=1e-12)
    if name != "ties":      # without ties the reference's default (unstable) sort gives the same answer
        np.testing.assert_array_equal(r["cmc"], g[f"{name}/cmc"])
        np.testing.assert_allclose(r["mAP"], g[f"{name}/mAP"], rtol=1e-12)
    cmc_c, map_c = MO.eval_market1501_c(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"], c["max_rank"])
    np.testing.assert_array_equal(cmc_c, g[f"{name}/cy_cmc_stable"])     # C twin == Cython reference, bit for bit
    assert np.float32(map_c) == np.float32(g[f"{name}/cy_mAP_stable"])
    np.testing.assert_allclose(cmc_c, r["cmc"], atol=1e-6)
    np.testing.assert_allclose(map_c, r["mAP"], rtol=1e-5)


@pytest.mark.parametrize("name", GI.MODEL_CASES)
def test_model_oracle_matches_reference(golden, name):
    g = golden("model_" + name)
    c = GI.model_case(name)
    x = torch.from_numpy(c["images"])
    for mode in ("train", "eval"):
        with torch.no_grad():
            emb, vis, ids, pix, feats, masks = MDL.bpbreid_forward(c["state_dict"], x, c["backbone"], c["K"],
                                                                   training=(mode == "train"))
        for k, v in emb.items():
            ref = g[f"{mode}/emb/{k}"]
            err = np.abs(v.numpy() - ref).max() / (np.abs(ref).max() + 1e-12)
            assert err < 2e-3, (mode, k, err)
        for k, v in ids.items():
            ref = g[f"{mode}/ids/{k}"]
            err = np.abs(v.numpy() - ref).max() / (np.abs(ref).max() + 1e-12)
            assert err < 2e-3, (mode, k, err)
        for k, v in vis.items():
            np.testing.assert_array_equal(v.numpy(), g[f"{mode}/vis/{k}"])
        np.testing.assert_allclose(pix.numpy(), g[f"{mode}/pix"], rtol=1e-3, atol=1e-4)
        if mode == "train":
            ne = {k: v.numpy() for k, v in emb.items()}
            nv = {k: v.numpy() for k, v in vis.items()}
            ni = {k: v.numpy() for k, v in ids.items()}
            loss, _ = LO.gilt(ne, nv, ni, c["pids"])
            np.testing.assert_allclose(loss, g["train/gilt_loss"], rtol=1e-4)


def oracle_grads(c, bpa_weight=0.35, quant=None, weights=None):
    """Gradients of GiLt (+ bpa_weight * body-part attention loss) w.r.t. every floating tensor of the state dict, by
    torch autograd over the oracle restatements (model_oracle + loss_oracle_torch)."""
    from oracle import loss_oracle_torch as LT
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
          for k, v in c["state_dict"].items()}
    emb, vis, ids, pix, feats, masks = MDL.bpbreid_forward(sd, torch.from_numpy(c["images"]), c["backbone"], c["K"],
                                                           training=True, quant=quant)
    pids = torch.from_numpy(c["pids"])
    loss = LT.gilt(emb, ids, pids, weights)
    if bpa_weight > 0:
        loss = loss + bpa_weight * LT.body_part_attention(pix, torch.from_numpy(c["masks"]))
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items() if isinstance(v, torch.Tensor) and v.grad is not None}


def test_torch_loss_oracle_matches_numpy_oracle_and_reference(golden):
    from oracle import loss_oracle_torch as LT
    g = golden("losses")
    for name in ("default", "market", "dups_singletons", "soft_margin"):
        c = GI.loss_case(name)
        emb = {k: torch.from_numpy(v) for k, v in c["emb"].items()}
        ids = {k: torch.from_numpy(v) for k, v in c["ids"].items()}
        loss = LT.gilt(emb, ids, torch.from_numpy(c["pids"]), c["weights"], c["margin"])
        np.testing.assert_allclose(float(loss), g[f"{name}/loss"], rtol=2e-5)
    lg, tm = torch.from_numpy(g["bpa/logits"]), torch.from_numpy(g["bpa/masks"])
    np.testing.assert_allclose(float(LT.body_part_attention(lg, tm)), g["bpa/loss"], rtol=1e-5)


def test_oracle_autograd_matches_reference_backward(golden):
    """The checker of the analytic backward kernels is pinned to the unmodified reference's loss.backward()."""
    g = golden("grads_hrnet32_k5_ctr")
    c = GI.model_case("hrnet32_k5_ctr")
    loss, grads = oracle_grads(c)
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-4)
    names = [k[5:] for k in g.files if k.startswith("norm/")]
    assert len(names) > 900
    worst = 0.0
    for k in names:
        ref = float(g["norm/" + k])
        got = float(grads[k].double().norm())
        if ref < 1e-4:                               # Linear biases in front of a train-mode BN: the true gradient is 0
            assert got < 1e-4, (k, got)
            continue
        worst = max(worst, abs(got - ref) / ref)
    assert worst < 1e-3, worst                       # fp32 re-association through ~300 layers (measured ~3e-6)
    for k in [f[5:] for f in g.files if f.startswith("full/")]:
        ref = g["full/" + k]
        if np.abs(ref).max() < 1e-4:
            continue
        err = np.abs(grads[k].numpy() - ref).max() / (np.abs(ref).max() + 1e-12)
        assert err < 2e-2, (k, err)


@pytest.mark.parametrize("name", list(GI.RERANK_CASES))
def test_rerank_oracle_reproduces_reference_bit_for_bit(golden, name):
    """oracle/rerank_oracle.py vs the unmodified torchreid/utils/rerank.py output stored in tests/golden/rerank.npz."""
    from oracle import rerank_oracle as RO
    c = GI.rerank_case(name)
    out = RO.re_ranking(c["q_g"], c["q_q"], c["g_g"], c["k1"], c["k2"], c["lam"])
    np.testing.assert_array_equal(out, golden("rerank")[f"{name}/final"])


def test_input_oracle_resize_is_bit_exact_against_cv2():
    """oracle/input_oracle.py restates cv2.resize (INTER_LINEAR on uint8, INTER_NEAREST index rule): pinned against cv2
    itself (opencv-python 4.13 in this image; skipped where cv2 is absent)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import input_oracle as IO
    rs = np.random.RandomState(0)
    for (h, w, hh, ww) in [(128, 64, 256, 128), (128, 64, 384, 128), (200, 77, 256, 128), (51, 23, 256, 128),
                           (37, 90, 64, 32), (300, 150, 256, 128), (500, 311, 384, 128), (256, 128, 256, 128),
                           (1, 64, 1, 128), (64, 1, 128, 1), (255, 129, 256, 128)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        np.testing.assert_array_equal(IO.resize_linear_u8(img, hh, ww), cv2.resize(img, (ww, hh), interpolation=cv2.INTER_LINEAR))
        m = rs.rand(h, w, 3).astype(np.float32)
        yi, xi = IO.nearest_index_cv2(hh, h), IO.nearest_index_cv2(ww, w)
        np.testing.assert_array_equal(m[yi][:, xi], cv2.resize(m, (ww, hh), interpolation=cv2.INTER_NEAREST).reshape(hh, ww, 3))


@pytest.mark.parametrize("name", list(GI.INPUT_CASES))
def test_input_oracle_reproduces_reference_transforms(golden, name):
    """oracle/input_oracle.py vs tests/golden/inputs.npz (the reference's mask transform classes + cv2, composed in
    build_transforms' order by oracle/make_golden.py inputs), and the product's restated group tables vs the reference's."""
    from bpbreid_b200 import data
    from oracle import input_oracle as IO
    g = golden("inputs")
    c = GI.input_case(name)
    np.testing.assert_array_equal(data.group_table(c["preprocess"]), g["groups/" + c["preprocess"]])
    groups = [np.nonzero(r)[0].tolist() for r in g["groups/" + c["preprocess"]]]
    for i, (img, f, p) in enumerate(zip(c["images"], c["fields"], c["params"])):
        x = IO.preprocess_image(img, c["height"], c["width"], (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), p["flip"], p["crop"], p["erase"])
        m = IO.preprocess_masks(f, c["height"], c["width"], groups, 4, c["strategy"], c["threshold"], c["softmax_weight"],
                                p["flip"], p["crop"], p["erase"])
        np.testing.assert_array_equal(x, g[f"{name}/images"][i])
        np.testing.assert_array_equal(m, g[f"{name}/masks"][i])
