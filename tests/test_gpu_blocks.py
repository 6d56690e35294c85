"""GPU (B200): TEACHER-FORCED per-block parity of the real networks in TRAINING mode (batch-statistics BN).

Why not an end-to-end comparison?  A randomly initialised HRNet / ResNet with batch-statistics BN is chaotic: two CPU
bf16 emulations that differ only in fp32 summation order end up 18 % apart at the feature map and the relative error
of *any* bf16 evaluation against fp32 grows ~1.3x per residual block (measured).  So every block of
the actual network (every conv / BN / residual / fuse / upsample kernel at its real shape) is checked on its own:
the CUDA plan records each block's input and output; the oracle's bf16-rounding variant of the same block is
evaluated on the CUDA input and must reproduce the CUDA output.  End-to-end parity is asserted in eval mode
(tests/test_gpu_path.py), where the network is well conditioned."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import model_oracle as MDL
from tests import golden_inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nchw(a):
    return a.t.float().permute(0, 3, 1, 2).cpu()


def rel(got, ref):
    return float((got - ref).norm() / (ref.norm() + 1e-12))


def run_traced(name, n=8):
    c = GI.model_case(name, n=n)
    m = c["model"]
    m.load_state_dict(c["state_dict"])
    m = m.to(DEV).train()
    m.debug_trace = True
    x = torch.from_numpy(c["images"]).to(DEV)
    with torch.no_grad():
        out = m(x)
    torch.cuda.synchronize()
    plan = next(iter(m._plans.values()))
    return c, m, plan, out


@pytest.mark.parametrize("name", GI.MODEL_CASES)
def test_every_block_matches_bf16_oracle_on_its_own_input(name):
    c, m, plan, out = run_traced(name)
    sd = c["state_dict"]
    ctx = MDL.Ctx(True, MDL.bf16_round)
    pre = MDL.BACKBONE_PREFIX
    fastreid = c["backbone"] == "fastreid_resnet_ibn"
    worst, checked = {}, 0
    for e in plan.trace:
        kind, names = e["kind"], e["names"]
        with torch.no_grad():
            if kind == "stem":
                x = MDL.bf16_round(torch.from_numpy(c["images"]))
                p_bn = MDL._SD(sd, names[1] + ".")
                ref = [MDL._conv_bn(ctx, x, sd[names[0] + ".weight"], p_bn, e["stride"], e["ksize"] // 2, relu=True)]
            elif kind == "conv_bn_relu":
                ref = [MDL._conv_bn(ctx, nchw(e["ins"][0]), sd[names[0] + ".weight"], MDL._SD(sd, names[1] + "."),
                                    e["stride"], e["ksize"] // 2, relu=True)]
            elif kind == "basic":
                ref = [MDL._basic_block(ctx, nchw(e["ins"][0]), MDL._SD(sd, names[0] + "."))]
            elif kind == "bottleneck":
                fn = MDL._res_bottleneck if not c["backbone"].startswith("hrnet") else MDL._bottleneck
                ref = [fn(ctx, nchw(e["ins"][0]), MDL._SD(sd, names[0] + "."), e["stride"])]
            elif kind == "fuse":
                ref = MDL._hr_fuse(ctx, [nchw(a) for a in e["ins"]], MDL._SD(sd, names[0] + "."), len(e["ins"]))
            elif kind == "maxpool":
                ref = [F.max_pool2d(nchw(e["ins"][0]), 3, 2, e["pad"], ceil_mode=bool(e["ceil"]))]
            elif kind == "upcat":
                br = [nchw(a) for a in e["ins"]]
                h, w = br[0].shape[2:]
                ref = [torch.cat([br[0]] + [ctx.q(F.interpolate(b, size=(h, w), mode="bilinear", align_corners=True))
                                            for b in br[1:]], 1)]
            else:
                raise AssertionError(kind)
        for got, r in zip(e["outs"], ref):
            err = rel(nchw(got), r)
            worst[kind] = max(worst.get(kind, 0.0), err)
            checked += 1
            # bf16 output rounding (2^-9) + one-ulp flips of intermediate bf16 values inside the block
            assert err < 1.5e-2, (kind, names, err)
    assert pre  # (prefix constant used by the head check below)
    assert checked > (100 if name == "hrnet32_k5" else 15)
    print(name, "blocks checked:", checked, "worst rel err per kind:", {k: round(v, 5) for k, v in worst.items()})
    # ---- head, teacher-forced on the CUDA feature map (fp32 math on both sides; tf32 Linear layers) ----
    emb, vis, ids, pix, feats, masks = out
    up = [e for e in plan.trace if e["kind"] == "upcat"]
    if up:      # HRNet: the head consumes the branch maps directly (fp32 interpolation, nothing rounded to bf16)
        br = [nchw(a) for a in up[0]["ins"]]
        hh, ww = br[0].shape[2:]
        f32 = torch.cat([br[0]] + [F.interpolate(b, size=(hh, ww), mode="bilinear", align_corners=True) for b in br[1:]], 1)
    else:
        f32 = feats.float().cpu()
    with torch.no_grad():
        o = MDL.bpbreid_head(MDL.Ctx(True, None), f32, sd, c["K"])
    np.testing.assert_allclose(pix.cpu().numpy(), o[3].numpy(), rtol=0, atol=2e-3 * float(o[3].abs().max()) + 1e-5)
    for k in ("globl", "foreg", "backg", "parts", "conct", "bn_globl", "bn_foreg", "bn_backg", "bn_parts", "bn_conct"):
        e_ = rel(emb[k].float().cpu(), o[0][k])
        assert e_ < 1e-2, (k, e_)              # BN1d over N=8 rows amplifies the tf32 (2^-11) input rounding
    for k in ("globl", "foreg", "backg", "parts", "conct"):
        e_ = rel(ids[k].float().cpu(), o[2][k])
        assert e_ < 1e-2, (k, e_)
    for k in ("backg", "foreg", "parts"):
        assert (vis[k].cpu() == o[1][k]).float().mean() > 0.95, k
    for k in ("backg", "foreg", "parts"):
        assert rel(masks[k].float().cpu(), o[5][k]) < 5e-3, k


def test_running_statistics_side_effect_matches_torch():
    """Train-mode forward must update BN running stats exactly like nn.BatchNorm2d (momentum 0.1, unbiased var)."""
    c, m, plan, out = run_traced("resnet50_k5", n=4)
    sd0 = c["state_dict"]
    x = MDL.bf16_round(torch.from_numpy(c["images"]))
    w = MDL.bf16_round(sd0[MDL.BACKBONE_PREFIX + "conv1.weight"])
    y = MDL.bf16_round(F.conv2d(x, w, None, 2, 3))
    rm = sd0[MDL.BACKBONE_PREFIX + "bn1.running_mean"].clone()
    rv = sd0[MDL.BACKBONE_PREFIX + "bn1.running_var"].clone()
    F.batch_norm(y, rm, rv, None, None, True, 0.1, 1e-5)
    bn = m.backbone_appearance_feature_extractor.bn1
    assert rel(bn.running_mean.cpu(), rm) < 1e-3 and rel(bn.running_var.cpu(), rv) < 1e-3
