"""GPU (B200): parity of the reference-facing API (bpbreid_b200.models / losses / metrics) against the CPU oracle
and the golden vectors minted from the reference, plus size-independent properties at BASELINE.json's full
evaluation size.  Everything goes through the C ABI of libbpbreid_sm100.so."""
import numpy as np
import pytest
import torch

from bpbreid_b200 import losses, metrics
from oracle import loss_oracle as LO
from oracle import metrics_oracle as MO
from oracle import model_oracle as MDL
from tests import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(d):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in d.items()}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


# ------------------------------------------------------------------------------------------ GiLt loss
@pytest.mark.parametrize("name", GI.LOSS_CASES)
def test_gilt_loss_matches_reference_golden_and_oracle(golden, name):
    g = golden("losses")
    c = GI.loss_case(name)
    crit = losses.GiLtLoss(c["weights"], use_visibility_scores=c["use_vis"], triplet_margin=c["margin"],
                           loss_name="part_averaged_triplet_loss", use_gpu=True, writer=None)
    loss, summary = crit(cu(c["emb"]), cu(c["vis"]), cu(c["ids"]), torch.from_numpy(c["pids"]).to(DEV))
    torch.cuda.synchronize()
    o_loss, o_sum = LO.gilt(c["emb"], c["vis"], c["ids"], c["pids"], c["weights"], c["use_vis"], c["margin"])
    assert abs(float(loss) - float(g[f"{name}/loss"])) <= 1e-3 * abs(float(g[f"{name}/loss"]))     # reference
    assert abs(float(loss) - float(o_loss)) <= 1e-3 * abs(float(o_loss))                             # oracle
    for key, info in summary.items():
        for kk, vv in info.items():
            ref = float(g[f"{name}/{key}/{kk}"])
            assert abs(float(vv) - ref) <= 1e-3 * max(abs(ref), 1e-3), (key, kk, float(vv), ref)
    # mining internals: hardest positive / negative indices are integer results -> exact vs the oracle
    tl = crit.part_triplet_loss.last
    e = c["emb"]["parts"].transpose(1, 0, 2)
    pd = LO.part_pairwise_dist(e)
    if c["use_vis"]:
        v = c["vis"]["parts"].T
        m = (v[:, None, :] & v[:, :, None]) if v.dtype == bool else np.sqrt(v[:, None, :] * v[:, :, None])
        dist = LO.masked_mean(pd, m)
    else:
        dist = pd.mean(0, dtype=np.float32)
    d_gpu = tl["dist"].cpu().numpy()
    assert np.abs(d_gpu - dist).max() < 2e-3          # sqrt near duplicates
    assert ((d_gpu == -1) == (dist == -1)).all()      # sentinel pattern identical


def test_triplet_no_valid_triplet_returns_none_with_warning():
    emb = torch.randn(6, 3, 16, device=DEV)
    pids = torch.arange(6, device=DEV)                # all singletons -> no positive anywhere
    res = losses.PartAveragedTripletLoss(0.3)(emb, pids)
    with pytest.warns(UserWarning, match="no valid triplets"):
        assert res.materialize() is None


def test_body_part_attention_loss(golden):
    g = golden("losses")
    crit = losses.BodyPartAttentionLoss("cl", 0.1)
    loss, summary = crit.forward_from_masks(torch.from_numpy(g["bpa/logits"]).to(DEV), torch.from_numpy(g["bpa/masks"]).to(DEV))
    assert abs(float(loss) - float(g["bpa/loss"])) < 1e-3 * float(g["bpa/loss"])
    assert abs(float(summary["pixls"]["a"]) - float(g["bpa/acc"])) < 1e-6


# ------------------------------------------------------------------------------------------ distance
@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("name", GI.DIST_CASES)
def test_part_distance_matches_reference_golden(golden, name, impl):
    g = golden("distance")
    c = GI.distance_case(name)
    t = lambda a: None if a is None else torch.from_numpy(a).to(DEV)  # noqa: E731
    d, pd = metrics.part_distance_device(t(c["qf"]), t(c["gf"]), t(c["qv"]), t(c["gv"]), c["strat"], c["metric"], impl)
    torch.cuda.synchronize()
    d, pd = d.cpu().numpy(), pd.cpu().numpy()
    rd, rpd = g[f"{name}/dist"], g[f"{name}/part_dist"]
    # sentinel (max+1) pattern is an exact property; values within 1e-3 (fp32) -- tf32 dot products keep 10 mantissa bits
    tol = 2e-3 if impl == 0 else 2e-3
    assert np.abs(pd - rpd).max() < tol * max(1.0, np.abs(rpd).max()), np.abs(pd - rpd).max()
    assert np.abs(d - rd).max() < tol * max(1.0, np.abs(rd).max()), np.abs(d - rd).max()
    if c["qv"] is not None and c["qv"].dtype == bool:
        inv = ~(c["qv"].T[:, :, None] & c["gv"].T[:, None, :])
        mv = pd[inv]
        assert (mv == mv.flat[0]).all() if mv.size else True          # one common max+1 value
        assert (pd[~inv] < mv.flat[0]).all() if mv.size else True


def test_reference_signature_returns_cpu_tensors(golden):
    c = GI.distance_case("bool_mean")
    d, pd = metrics.compute_distance_matrix_using_bp_features(
        torch.from_numpy(c["qf"]), torch.from_numpy(c["gf"]), torch.from_numpy(c["qv"]), torch.from_numpy(c["gv"]),
        "mean", 500, True, "euclidean")
    assert d.device.type == "cpu" and pd.device.type == "cpu" and d.dtype == torch.float32
    g = golden("distance")
    assert np.abs(d.numpy() - g["bool_mean/dist"]).max() < 5e-3
    with pytest.raises(ValueError):
        metrics.compute_distance_matrix_using_bp_features(torch.zeros(2, 1, 32), torch.zeros(2, 1, 32), metric="l1")
    with pytest.raises(ValueError):
        metrics.compute_distance_matrix_using_bp_features(torch.zeros(2, 1, 32), torch.zeros(2, 1, 32),
                                                          dist_combine_strat="median")


# ------------------------------------------------------------------------------------------ ranking
@pytest.mark.parametrize("name", GI.RANK_CASES)
def test_rank_bit_exact_vs_reference_stable(golden, name):
    g = golden("rank")
    c = GI.rank_case(name)
    r = metrics.evaluate_rank(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"], c["max_rank"])
    np.testing.assert_array_equal(r["cmc"], g[f"{name}/cmc_stable"])                 # reference numpy path, stable sort
    assert abs(r["mAP"] - float(g[f"{name}/mAP_stable"])) < 1e-12
    cmc, m_ap = metrics.evaluate_cy(c["dist"], c["q_pids"], c["g_pids"], c["q_camids"], c["g_camids"], c["max_rank"])
    np.testing.assert_array_equal(cmc, g[f"{name}/cy_cmc_stable"])                   # reference Cython module
    assert np.float32(m_ap) == np.float32(g[f"{name}/cy_mAP_stable"])
    idx = metrics.argsort_rows(torch.from_numpy(c["dist"]).to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(idx, np.argsort(c["dist"], axis=1, kind="stable"))  # ranking indices: bit-exact


def test_rank_asserts_when_no_query_is_valid():
    d = np.random.RandomState(0).rand(4, 9).astype(np.float32)
    with pytest.raises(AssertionError, match="do not appear in gallery"):
        metrics.evaluate_rank(d, np.arange(4) + 100, np.arange(9), np.zeros(4), np.ones(9))
    with pytest.raises(ValueError):
        metrics.evaluate_rank(d, np.arange(4), np.arange(9), np.zeros(4), np.ones(9), eval_metric="zz")


def test_eval_pipeline_medium_scale_vs_oracle():
    """1024 q x 4096 g x 6 parts: distance (tcgen05) -> ranking, against the CPU oracle end to end."""
    rs = np.random.RandomState(0)
    nq, ng, p, d = 1024, 4096, 6, 512
    qf = MO.l2_normalize(rs.randn(nq, p, d).astype(np.float32))
    gf = MO.l2_normalize(rs.randn(ng, p, d).astype(np.float32))
    qv, gv = rs.rand(nq, p) > 0.15, rs.rand(ng, p) > 0.15
    qp, gp = rs.randint(0, 400, nq), rs.randint(0, 400, ng)
    qc, gc = rs.randint(0, 6, nq), rs.randint(0, 6, ng)
    dist, pd = metrics.part_distance_device(torch.from_numpy(qf).to(DEV), torch.from_numpy(gf).to(DEV),
                                            torch.from_numpy(qv).to(DEV), torch.from_numpy(gv).to(DEV))
    od, opd = MO.part_distance(qf, gf, qv, gv)
    assert np.abs(dist.cpu().numpy() - od).max() < 2e-3
    assert np.abs(pd.cpu().numpy() - opd).max() < 2e-3
    # ranking on the IDENTICAL fp32 matrix: exact agreement with the oracle (stable order)
    dn = dist.cpu().numpy()
    r = metrics.evaluate_rank(dn, qp, gp, qc, gc, 50)
    o = MO.eval_market1501(dn, qp, gp, qc, gc, 50)
    np.testing.assert_array_equal(r["cmc"], o["cmc"])
    assert abs(r["mAP"] - o["mAP"]) < 1e-12
    cmc_c, map_c = MO.eval_market1501_c(dn, qp, gp, qc, gc, 50)
    cmc, m32 = metrics.evaluate_cy(dn, qp, gp, qc, gc, 50)
    np.testing.assert_array_equal(cmc, cmc_c)
    assert np.float32(m32) == np.float32(map_c)
    sub = dist[:64].contiguous()
    np.testing.assert_array_equal(metrics.argsort_rows(sub).cpu().numpy(), np.argsort(dn[:64], axis=1, kind="stable"))


def test_full_size_config5_properties():
    """BASELINE.json config 5 (4096 q x 16384 g, K=5 -> P=6): properties that need no CPU reference."""
    gq = torch.Generator(device=DEV).manual_seed(0)
    nq, ng, p, d = 4096, 16384, 6, 512
    qf = torch.nn.functional.normalize(torch.randn(nq, p, d, device=DEV, generator=gq), dim=-1)
    gf = torch.nn.functional.normalize(torch.randn(ng, p, d, device=DEV, generator=gq), dim=-1)
    gf[:nq] = qf                                           # plant every query in the gallery
    qv = torch.rand(nq, p, device=DEV, generator=gq) > 0.15
    gv = torch.rand(ng, p, device=DEV, generator=gq) > 0.15
    qv[:, 0] = True
    gv[:nq] = qv                                           # same visibility for the planted copies
    dist, pd = metrics.part_distance_device(qf, gf, qv, gv)
    diag = dist[torch.arange(nq), torch.arange(nq)]
    assert float(diag.abs().max()) < 5e-3                  # d(x, x) ~ 0 (measured 2.3e-3): norms come from the tf32-truncated inputs
    assert int((dist.argmin(1) == torch.arange(nq, device=DEV)).sum()) == nq
    inv = ~(qv.t()[:, :, None] & gv.t()[:, None, :])
    sentinel = pd.max()
    assert bool((pd[inv] == sentinel).all()) and bool((pd[~inv] <= 2.001).all())
    # visibility-masked mean recomputed from the per-part matrix (checksum of checksums)
    w = (~inv).float()
    rec = (torch.where(inv, torch.zeros_like(pd), pd)).sum(0) / w.sum(0).clamp(min=1)
    ok = w.sum(0) > 0
    assert float((rec[ok] - dist[ok]).abs().max()) < 1e-4
    # ranking: planted copy with another camera must be rank-0 for every query -> cmc == 1, mAP == 1
    pid = torch.arange(ng, device=DEV)
    pid[:nq] = torch.arange(nq, device=DEV)
    pid[nq:] += 10 ** 6
    res = metrics.rank_device(dist, torch.arange(nq), pid.cpu(), torch.zeros(nq), torch.ones(ng), 50)
    assert int(res["num_valid"]) == nq and float(res["cmc"][0]) == 1.0 and abs(float(res["mAP"]) - 1.0) < 1e-12


# ------------------------------------------------------------------------------------------ full model
def _run_model(name, training, conv_impl="tc"):
    c = GI.model_case(name)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to(DEV)
    m.train(training)
    m.conv_impl = conv_impl
    x = torch.from_numpy(c["images"]).to(DEV)
    with torch.no_grad():
        out = m(x)
    torch.cuda.synchronize()
    return c, m, out


@pytest.mark.parametrize("name", GI.MODEL_CASES)
def test_model_eval_forward_matches_reference_golden(golden, name):
    """End-to-end parity against the REFERENCE's own outputs in eval mode (running-statistics BN: well conditioned).
    Training-mode end-to-end comparisons are chaotic for a random-init network: training mode is
    covered block by block in tests/test_gpu_blocks.py instead."""
    mode = "eval"
    g = golden("model_" + name)
    c, m, (emb, vis, ids, pix, feats, masks) = _run_model(name, False)
    errs = {}
    for k in ("globl", "foreg", "backg", "parts", "conct", "bn_globl", "bn_foreg", "bn_parts", "bn_conct"):
        ref = g[f"{mode}/emb/{k}"]
        got = emb[k].float().cpu().numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        errs[k] = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12)
    # bf16 activations through ~100 layers: this is the storage-format noise floor -- the CPU bf16-emulating oracle
    # deviates from the fp32 reference by the same amount (parts 1.35e-2 on this case); see the test below for the
    # tight comparison against that emulation
    # ill-conditioned exception, 384x128 case only: the background embedding there is a small difference of large terms
    # (norm 10x below the other embeddings); the CPU bf16-emulating oracle itself is 1.04e-1 away from the fp32
    # reference on it (5.4e-2 measured for the CUDA path), while every other embedding stays at ~1e-2
    loose = {"backg": 0.12} if name == "hrnet32_k5_384" else {}
    assert all(v < loose.get(k, 3e-2) for k, v in errs.items()), errs
    for k in ("globl", "foreg", "parts", "conct"):
        ref = g[f"{mode}/ids/{k}"]
        got = ids[k].float().cpu().numpy()
        assert got.shape == ref.shape
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 2e-2, k
    assert feats.shape[1] == (1920 if c["backbone"] == "hrnet32" else 2048)
    fm = feats.float().mean(dim=(2, 3)).cpu().numpy()
    assert np.linalg.norm(fm - g[f"{mode}/feat_mean"]) / np.linalg.norm(g[f"{mode}/feat_mean"]) < 2e-2
    pref = g[f"{mode}/pix"]
    # (2.3e-2 measured on the 96x32 map of the 384x128 case, 1.2-1.6e-2 on the 64x32 / 16x8 maps)
    assert np.linalg.norm(pix.cpu().numpy() - pref) / np.linalg.norm(pref) < (3e-2 if name == "hrnet32_k5_384" else 2e-2)
    agree = (vis["parts"].cpu().numpy() == g[f"{mode}/vis/parts"]).mean()
    assert agree > 0.9, agree


@pytest.mark.parametrize("name", GI.MODEL_CASES)
def test_model_eval_forward_matches_bf16_emulating_oracle(name):
    """Same rounding points on the CPU (oracle quant hook): isolates kernel bugs from bf16 storage noise."""
    c, m, (emb, vis, ids, pix, feats, masks) = _run_model(name, False)
    with torch.no_grad():
        o = MDL.bpbreid_forward(c["state_dict"], torch.from_numpy(c["images"]), c["backbone"], c["K"], training=False,
                                quant=MDL.bf16_round)
    for k in ("globl", "foreg", "backg", "parts", "bn_parts"):
        ref = o[0][k].numpy()
        got = emb[k].float().cpu().numpy()
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        # the head reads the fp32-interpolated branch maps, the emulation a bf16-rounded concat: ~1.6e-2 measured
        # (384x128 background embedding: ill-conditioned, see the test above)
        assert err < (0.12 if (name == "hrnet32_k5_384" and k == "backg") else 2.5e-2), (k, err)


def test_training_forward_and_loss_are_finite_and_consistent():
    """Training-mode smoke of the public API: shapes / dtypes of the 6-tuple, finite GiLt loss, and the loss equals the
    oracle's GiLt evaluated on the SAME (CUDA-produced) embeddings."""
    from oracle import loss_oracle as LO2
    c, m, (emb, vis, ids, pix, feats, masks) = _run_model("hrnet32_k5", True)
    n, K = 4, 5
    assert emb["parts"].shape == (n, K, 512) and emb["conct"].shape == (n, K * 512) and ids["parts"].shape == (n, K, 751)
    assert vis["parts"].dtype is torch.bool and vis["parts"].shape == (n, K) and vis["globl"].all()
    assert pix.shape == (n, K + 1, 64, 32) and feats.shape == (n, 1920, 64, 32) and masks["parts"].shape == (n, K, 64, 32)
    crit = losses.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, use_gpu=True)
    pids = torch.from_numpy(c["pids"]).to(DEV)
    loss, summary = crit(emb, vis, ids, pids)
    assert torch.isfinite(loss)
    ne = {k: v.float().cpu().numpy() for k, v in emb.items()}
    nv = {k: v.cpu().numpy() for k, v in vis.items()}
    ni = {k: v.float().cpu().numpy() for k, v in ids.items()}
    o_loss, _ = LO2.gilt(ne, nv, ni, c["pids"])
    assert abs(float(loss) - float(o_loss)) < 1e-3 * abs(float(o_loss))


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
def test_stream_lane_schedule_is_race_free(monkeypatch, graph):
    """The plan records HRNet branches / fuse outputs / head groups on four stream lanes (fork/join).  A missing
    dependency or a buffer reused across lanes would corrupt activations, so the multi-stream schedule must reproduce
    itself and the single-stream schedule.  Eval mode (well conditioned; the only run-to-run freedom left is the order of
    the shared-memory float atomics in the pixel head, ~1e-7)."""
    def run(single):
        monkeypatch.setenv("BPB_SINGLE_STREAM", "1" if single else "0")
        c = GI.model_case("hrnet32_k5")
        m = c["model"]
        m.load_state_dict(c["state_dict"])
        m = m.to(DEV)
        m.train(False)
        m.use_cuda_graph = graph
        x = torch.from_numpy(c["images"]).to(DEV)
        outs = []
        with torch.no_grad():
            for _ in range(3):
                emb, vis, ids, pix, _, masks = m(x)
                torch.cuda.synchronize()
                outs.append([t.float().clone() for t in (emb["parts"], emb["bn_globl"], ids["conct"], pix)])
        pl = next(iter(m._plans.values()))
        return outs, {op[3] for op in pl.ops}

    multi, lanes_m = run(False)
    single, lanes_s = run(True)
    assert lanes_m == {0, 1, 2, 3} and lanes_s == {0}

    def close(a, b):
        # run-to-run freedom (atomics order) is ~5e-5 here; a race would corrupt whole activations (O(1) errors)
        return float((a - b).abs().max()) <= 1e-3 * float(b.abs().max()) + 1e-7

    for other in (multi[1], multi[2], single[0]):
        for a, b in zip(multi[0], other):
            assert close(a, b)


def test_tc_and_simt_conv_paths_agree_in_eval_mode():
    _, _, out_tc = _run_model("hrnet32_k5", False, "tc")
    _, _, out_ref = _run_model("hrnet32_k5", False, "ref")
    a, b = out_tc[0]["parts"].float().cpu().numpy(), out_ref[0]["parts"].float().cpu().numpy()
    # two bf16 pipelines that differ only in fp32 accumulation order: ~1.3e-2 after ~100 layers (measured)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 2.5e-2


def test_map_parity_on_synthetic_gallery_from_model_embeddings(golden):
    """north_star: ranking / mAP on a held synthetic gallery.  Test embeddings ['bn_foreg','parts'] of the eval-mode
    model (part_based_engine.py:365-387) -> L2 normalise -> part distance -> ranking, CUDA vs oracle end to end."""
    name = "resnet50_k5"
    c = GI.model_case(name, n=4)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to(DEV).eval()
    gimg = torch.Generator().manual_seed(77)
    imgs = torch.randn(48, 3, 256, 128, generator=gimg)
    feats_c, vis_c, feats_o, vis_o = [], [], [], []
    for i in range(0, 48, 8):
        x = imgs[i:i + 8]
        with torch.no_grad():
            emb, vis, *_ = m(x.to(DEV))
            e = torch.cat([emb["bn_foreg"].unsqueeze(1), emb["parts"]], 1).float().cpu()
            v = torch.cat([vis["foreg"].unsqueeze(1), vis["parts"]], 1).cpu()
            feats_c.append(e.clone()); vis_c.append(v.clone())
            o = MDL.bpbreid_forward(c["state_dict"], x, c["backbone"], c["K"], training=False)
            feats_o.append(torch.cat([o[0]["bn_foreg"].unsqueeze(1), o[0]["parts"]], 1))
            vis_o.append(torch.cat([o[1]["foreg"].unsqueeze(1), o[1]["parts"]], 1))
    fc, vc = torch.cat(feats_c), torch.cat(vis_c)
    fo, vo = torch.cat(feats_o), torch.cat(vis_o)
    assert float((fc - fo).norm() / fo.norm()) < 2e-2
    fcn = torch.nn.functional.normalize(fc, dim=-1)
    pids = np.arange(48) % 6
    cams = np.arange(48) // 24
    q, gal = slice(0, 16), slice(16, 48)
    # identical fp32 inputs -> identical distances (1e-3) and bit-identical ranking on identical distance matrices
    d_c, _ = metrics.part_distance_device(fcn[q].to(DEV), fcn[gal].to(DEV), vc[q].to(DEV), vc[gal].to(DEV))
    d_o, _ = MO.part_distance(fcn[q].numpy(), fcn[gal].numpy(), vc[q].numpy(), vc[gal].numpy())
    assert np.abs(d_c.cpu().numpy() - d_o).max() < 2e-3
    r_c = metrics.evaluate_rank(d_c.cpu().numpy(), pids[q], pids[gal], cams[q], cams[gal], 20)
    r_o = MO.eval_market1501(d_c.cpu().numpy(), pids[q], pids[gal], cams[q], cams[gal], 20)
    np.testing.assert_array_equal(r_c["cmc"], r_o["cmc"])
    assert abs(r_c["mAP"] - r_o["mAP"]) < 1e-12


def test_end_to_end_map_cuda_distance_and_rank_vs_reference_distance_and_rank():
    """Reference chain (fp32 distance oracle -> numpy ranking) against the CUDA chain (bpb_part_distance ->
    bpb_rank_market1501) on the SAME features: identity-clustered synthetic gallery, bool visibility.  The SIMT fp32
    distance (impl=1) must give identical CMC and mAP; the tcgen05 path feeds tf32-rounded operands (errors up to 2e-3
    on a distance of ~1), which may swap near-equal neighbours: its CMC / mAP are asserted to the measured bound and the
    number of differing rank positions is reported."""
    rs = np.random.RandomState(21)
    nid, p, d, nq, ng = 40, 6, 512, 128, 1024
    cen = rs.randn(nid, p, d).astype(np.float32)

    def feats(n):
        ids = rs.randint(0, nid, n)
        f = cen[ids] + 1.5 * rs.randn(n, p, d).astype(np.float32)
        return (f / np.linalg.norm(f, axis=-1, keepdims=True)).astype(np.float32), ids, rs.rand(n, p) > 0.2
    (qf, qid, qv), (gf, gid, gv) = feats(nq), feats(ng)
    qc, gc = rs.randint(0, 3, nq), rs.randint(0, 3, ng)
    d_ref, _ = MO.part_distance(qf, gf, qv, gv)
    r_ref = MO.eval_market1501(d_ref, qid, gid, qc, gc, 50)
    rep = []
    for impl, name in ((1, "simt fp32"), (0, "tcgen05 tf32")):
        d_c, _ = metrics.part_distance_device(torch.from_numpy(qf).to(DEV), torch.from_numpy(gf).to(DEV),
                                              torch.from_numpy(qv).to(DEV), torch.from_numpy(gv).to(DEV), impl=impl)
        r_c = metrics.evaluate_rank(d_c.cpu().numpy(), qid, gid, qc, gc, 50)
        swaps = int((np.argsort(d_c.cpu().numpy(), 1, kind="stable") != np.argsort(d_ref, 1, kind="stable")).sum())
        rep.append(f"{name}: max |d - d_ref| {np.abs(d_c.cpu().numpy() - d_ref).max():.2e}, differing rank positions "
                   f"{swaps} of {nq * ng}, |cmc diff| {np.abs(r_c['cmc'] - r_ref['cmc']).max():.2e}, "
                   f"mAP {r_c['mAP']:.6f} vs reference {r_ref['mAP']:.6f}")
        if impl == 1:
            np.testing.assert_allclose(r_c["cmc"], r_ref["cmc"], atol=1e-7)
            assert abs(r_c["mAP"] - r_ref["mAP"]) < 1e-6, rep
        else:
            assert np.abs(r_c["cmc"] - r_ref["cmc"]).max() <= 1.0 / nq + 1e-7 and abs(r_c["mAP"] - r_ref["mAP"]) < 1e-3, rep
    print("\n".join(rep))
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/map_e2e_report.txt", "w").write("\n".join(rep) + "\n")


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_masked_sentinel_value_with_ragged_query_tiles(metric, impl):
    """Nq not a multiple of the 128-row tile, non-negative un-normalised features, cosine and euclidean: the
    rows of the last query tile beyond Nq (TMA zero fill) must not leak into max(part_dist) + 1, the value every masked
    entry is rewritten to (distance.py:173-176)."""
    rs = np.random.RandomState(12)
    nq, ng, p, d = 200, 300, 6, 64
    qf = (rs.rand(nq, p, d) * 0.05).astype(np.float32)           # small norms: a leaked |g| or 1.0 would dominate
    gf = (rs.rand(ng, p, d) * 3.0).astype(np.float32)
    qv, gv = rs.rand(nq, p) > 0.3, rs.rand(ng, p) > 0.3
    d_o, p_o = MO.part_distance(qf, gf, qv, gv, "mean", metric)
    d_c, p_c = metrics.part_distance_device(torch.from_numpy(qf).to(DEV), torch.from_numpy(gf).to(DEV),
                                            torch.from_numpy(qv).to(DEV), torch.from_numpy(gv).to(DEV), "mean", metric, impl=impl)
    p_c, d_c = p_c.cpu().numpy(), d_c.cpu().numpy()
    masked = ~(qv.T[:, :, None] & gv.T[:, None, :])
    tol = 2e-3 if impl == 0 else 1e-5
    assert abs(float(p_c[masked].max()) - float(p_o[masked].max())) < tol * max(1.0, float(p_o.max()))   # the sentinel itself
    assert (p_c[masked] == p_c[masked].max()).all()
    np.testing.assert_allclose(p_c, p_o, rtol=tol, atol=tol * max(1.0, float(p_o.max())))
    np.testing.assert_allclose(d_c, d_o, rtol=tol, atol=tol * max(1.0, float(p_o.max())))


def test_argsort_rows_beyond_16384_entries():
    """Market-1501's gallery has 19 732 entries: the split-array sorter (keys + 16-bit indices) == np.argsort(stable),
    including massive tie groups."""
    rs = np.random.RandomState(2)
    d = rs.rand(6, 19732).astype(np.float32)
    d[1] = np.round(d[1] * 8) / 8
    d[2, :5000] = 2.0
    idx = metrics.argsort_rows(torch.from_numpy(d).to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(idx, np.argsort(d, axis=1, kind="stable"))


@pytest.mark.parametrize("name", list(GI.MASK_CASES))
def test_external_part_masks_match_reference_golden(golden, name):
    """External part masks (bpbreid.py:149-175): attention taken from the masks (learnable_attention_enabled False), and
    the 'hard' / 'soft' target segmentation of the learned attention at test time -- against the reference's outputs."""
    g = golden("model_" + name)
    c = GI.mask_case(name)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to("cuda").eval()
    x = torch.from_numpy(c["images"]).to("cuda")
    ext = torch.from_numpy(c["masks"]).to("cuda")
    with pytest.raises(AssertionError):
        m(x)                                                   # the reference asserts the masks are there
    with torch.no_grad():
        emb, vis, ids, pix, feats, masks = m(x, external_parts_masks=ext)
    assert (pix is not None) == bool(g["eval/has_pix"])
    errs = {}
    for k in ("globl", "foreg", "backg", "parts", "bn_foreg", "bn_parts"):
        ref = g[f"eval/emb/{k}"]
        got = emb[k].float().cpu().numpy()
        errs[k] = float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-12))
    for k in ("backg", "foreg", "parts"):
        ref = g[f"eval/masks/{k}"]
        got = masks[k].float().cpu().numpy()
        # mean error: with random-init weights in eval mode the pixel logits are huge (~1e7), i.e. the softmax is an
        # arg-max and a handful of pixels flip class under bf16 noise -- a max-norm over pixels would only see those
        errs["mask/" + k] = float(np.abs(got - ref).mean() / (np.abs(ref).mean() + 1e-12))
    print(name, errs)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/mask_cases_report.txt", "a") as fh:
        fh.write(f"{name}: {errs}\n")
    # masks: 2.5-3.9e-2 mean error measured = the few per cent of pixels whose arg-max flips (see above)
    assert all(v < (5e-2 if k.startswith("mask/") else 3e-2) for k, v in errs.items()), errs
    agree = (vis["parts"].cpu().numpy() == g["eval/vis/parts"]).mean()
    assert agree > 0.9, agree


@pytest.mark.parametrize("name", ["hrnet32_k5", "resnet50_k5", "hrnet32_k5_384"])
def test_inference_epilogue_fusion_is_bit_identical(name):
    """Inference: BatchNorm / residual / ReLU in the convolution's epilogue (bpb_conv_bn_act_bf16) must reproduce the
    conv -> bn_apply path bit for bit (the BatchNorm is applied to the bf16-rounded accumulator in both; projection
    skips and fuse-layer terms, which the two-pass path combines in fp32, keep the two-pass path)."""
    import os
    c = GI.model_case(name)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to("cuda").eval()
    x = torch.from_numpy(c["images"]).to("cuda")
    m.materialize_spatial_features = True
    res = {}
    for single in ("1", "0"):
        os.environ["BPB_SINGLE_STREAM"] = single
        for fuse in (True, False):
            m._plans = {}
            m.fuse_eval_epilogue = fuse
            runs = []
            for _ in range(2):
                with torch.no_grad():
                    emb, vis, ids, pix, feats, masks = m(x)
                runs.append((feats.clone(), {k: v.clone() for k, v in emb.items()}))
            plan = m.plan_for(x)
            names = [op[2] for op in plan.ops if op[0] is not None]
            res[(single, fuse)] = (runs, names)
    os.environ.pop("BPB_SINGLE_STREAM", None)
    rep = []
    for key, (runs, names) in res.items():
        rep.append((key, "self-consistent", torch.equal(runs[0][0], runs[1][0]), len(names), names.count("bpb_bn_apply")))
    base = res[("1", False)][0][0][0].float()
    for key, (runs, names) in res.items():
        d = (runs[0][0].float() - base).abs().max() / base.abs().max()
        rep.append((key, "feature-map max rel diff vs single-stream two-pass", float(d)))
    print(rep)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fusion_report.txt", "a") as fh:
        fh.write(f"{name}: {rep}\n")
    n1, n0 = res[("0", True)][1], res[("0", False)][1]
    assert "bpb_conv_bn_act_bf16" in n1 and "bpb_conv_bn_act_bf16" not in n0
    assert n1.count("bpb_bn_apply") < n0.count("bpb_bn_apply") // 4 and len(n1) < len(n0)
    for key, (runs, names) in res.items():
        assert torch.equal(runs[0][0], res[("1", False)][0][0][0]), key
