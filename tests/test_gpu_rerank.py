"""-m gpu: k-reciprocal re-ranking (csrc/rerank.cu through the C-ABI) against the reference's own output
(tests/golden/rerank.npz, minted by oracle/make_golden.py from torchreid/utils/rerank.py) and, stage by stage, against the
oracle restatement.  Integer stages (neighbour table, expanded k-reciprocal sets) bit-exact; floating-point output within
2e-6 absolute (values in [0, 1]; the only non-IEEE-identical operation is exp)."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import rerank_oracle as RO
from tests import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _run(c, stages=False):
    from bpbreid_b200 import rerank
    args = [torch.from_numpy(c[k]).to(DEV) for k in ("q_g", "q_q", "g_g")]
    return rerank.re_ranking_device(*args, k1=c["k1"], k2=c["k2"], lambda_value=c["lam"], return_stages=stages)


@pytest.mark.parametrize("name", list(GI.RERANK_CASES))
def test_rerank_matches_reference_golden(golden, name):
    c = GI.rerank_case(name)
    out, st = _run(c, True)
    torch.cuda.synchronize()
    ref = golden("rerank")[f"{name}/final"]
    _, ost = RO.re_ranking(c["q_g"], c["q_q"], c["g_g"], c["k1"], c["k2"], c["lam"], return_stages=True)
    a = c["q_g"].shape[0] + c["q_g"].shape[1]
    # dense normalised distance: IEEE multiply / divide only -> identical
    np.testing.assert_array_equal(st["orig"].cpu().numpy(), ost["orig"])
    # neighbour table (integer work): bit-exact
    kk = ost["rank"].shape[1]
    np.testing.assert_array_equal(st["rank"].cpu().numpy()[:, :kk], ost["rank"])
    # expanded k-reciprocal sets: identical as sets, weights within 1e-6
    vidx, vval, vcnt = (st[k].cpu().numpy() for k in ("vidx", "vval", "vcnt"))
    for i in range(a):
        got = np.sort(vidx[i, :vcnt[i]])
        np.testing.assert_array_equal(got, ost["sets"][i])
    # final distance vs the reference itself
    err = np.abs(out.cpu().numpy() - ref).max()
    print(f"rerank {name}: max abs err vs reference {err:.2e}")
    assert out.shape == ref.shape and err < 2e-6
    # ranking induced by the re-ranked distance: identical up to exact-tie groups of the reference
    assert (np.argsort(out.cpu().numpy(), 1, kind="stable") == np.argsort(ref, 1, kind="stable")).mean() > 0.98


def test_rerank_numpy_signature_and_engine_hook():
    """re_ranking(numpy, numpy, numpy) -> numpy float32, like torchreid.utils.re_ranking (part_based_engine.py:226)."""
    from bpbreid_b200 import rerank
    c = GI.rerank_case("clusters")
    out = rerank.re_ranking(c["q_g"], c["q_q"], c["g_g"])
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == c["q_g"].shape
    ref = RO.re_ranking(c["q_g"], c["q_q"], c["g_g"])
    assert np.abs(out - ref).max() < 2e-6
    with pytest.raises(ValueError):
        rerank.re_ranking(c["q_g"], c["q_q"][:-1], c["g_g"])


def test_rerank_medium_scale_vs_oracle_and_improves_map():
    """600 x 2400 clustered features: product vs oracle (1e-5), and re-ranking must not hurt mAP on clustered data."""
    from bpbreid_b200 import metrics, rerank
    rs = np.random.RandomState(5)
    nid, d = 150, 64
    cen = rs.randn(nid, d).astype(np.float32)

    def feats(n):
        ids = rs.randint(0, nid, n)
        f = cen[ids] + 0.9 * rs.randn(n, d).astype(np.float32)
        return torch.from_numpy(f / np.linalg.norm(f, axis=1, keepdims=True)).to(DEV), ids
    (qf, qid), (gf, gid) = feats(600), feats(2400)
    qg, qq, gg = torch.cdist(qf, gf), torch.cdist(qf, qf), torch.cdist(gf, gf)
    out = rerank.re_ranking_device(qg, qq, gg)
    torch.cuda.synchronize()
    ref = RO.re_ranking(qg.cpu().numpy(), qq.cpu().numpy(), gg.cpu().numpy())
    err = np.abs(out.cpu().numpy() - ref).max()
    print(f"rerank 600x2400: max abs err vs oracle {err:.2e}")
    assert err < 1e-5
    qc, gc = np.zeros(600, np.int64), np.ones(2400, np.int64)
    m0 = metrics.evaluate_rank(qg.cpu().numpy(), qid, gid, qc, gc, 50)["mAP"]
    m1 = metrics.evaluate_rank(out.cpu().numpy(), qid, gid, qc, gc, 50)["mAP"]
    print(f"mAP {m0:.4f} -> {m1:.4f} after re-ranking")
    assert m1 > m0


def test_rerank_market_scale_runs_and_is_sane():
    """Market-1501 scale (3368 x 15913, A = 19281: rows no longer fit the 16384-entry sorter, the dense reference needs
    three 1.5 GB matrices and minutes of Python loops): timing + range / finiteness properties."""
    from bpbreid_b200 import rerank
    g = torch.Generator(device=DEV).manual_seed(3)
    nq, ng, d = 3368, 15913, 128
    cen = torch.randn(751, d, device=DEV, generator=g)
    qf = torch.nn.functional.normalize(cen[torch.randint(0, 751, (nq,), device=DEV, generator=g)] + 0.8 * torch.randn(nq, d, device=DEV, generator=g), dim=1)
    gf = torch.nn.functional.normalize(cen[torch.randint(0, 751, (ng,), device=DEV, generator=g)] + 0.8 * torch.randn(ng, d, device=DEV, generator=g), dim=1)
    qg, qq, gg = torch.cdist(qf, gf), torch.cdist(qf, qf), torch.cdist(gf, gf)
    rerank.re_ranking_device(qg, qq, gg)
    torch.cuda.synchronize()
    t0 = time.time()
    out = rerank.re_ranking_device(qg, qq, gg)
    torch.cuda.synchronize()
    ms = (time.time() - t0) * 1e3
    msg = f"re-ranking {nq} x {ng} (k1=20, k2=6): {ms:.1f} ms on the device"
    print(msg)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/rerank_report.txt", "w").write(msg + "\n")
    assert torch.isfinite(out).all() and float(out.min()) >= 0 and float(out.max()) <= 1.0 + 1e-6
