"""GPU (B200): kernel-level parity of the sm_100a library, called through the C ABI.

Checkers: the CPU oracle restatements (oracle/), the committed golden vectors minted from the reference
(tests/golden/), and -- for the tensor-core convolution -- a plain fp32 PyTorch evaluation of the same bf16 operands.
Tolerances: bit-exact for integer / index results; fp32 kernels 1e-3 relative (north_star); bf16 storage rounding
2^-8 relative on outputs that are stored as bf16."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from bpbreid_b200 import _lib

pytestmark = pytest.mark.gpu

DEV = "cuda"


def S():
    return torch.cuda.current_stream().cuda_stream


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def pack_w(w):   # [co, ci, kh, kw] -> OHWI bf16
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()


CONV_SHAPES = [
    # n, h, w, cin, cout, k, s
    (4, 64, 32, 32, 32, 3, 1),     # HRNet branch 0 basic block (64-byte rows, SW64)
    (4, 32, 16, 64, 64, 3, 1),     # branch 1 (128-byte rows, SW128)
    (4, 16, 8, 128, 128, 3, 1),    # branch 2, one image per tile
    (8, 8, 4, 256, 256, 3, 1),     # branch 3, four images per tile
    (3, 8, 4, 256, 256, 3, 1),     # ragged: last tile has fewer images than the box
    (4, 64, 32, 64, 256, 1, 1),    # layer1 1x1 expand
    (4, 64, 32, 256, 64, 1, 1),    # layer1 1x1 reduce
    (4, 64, 32, 256, 64, 3, 2),    # transition1 stride-2 (TMA element strides)
    (4, 128, 64, 64, 64, 3, 2),    # stem conv2 stride-2
    (4, 64, 32, 32, 32, 3, 2),     # fuse chain stride-2 with 64-byte rows
    (2, 32, 16, 512, 1024, 1, 2),  # ResNet downsample 1x1 stride-2
    (2, 16, 8, 1024, 2048, 1, 1),  # ResNet layer4 expand
    (2, 16, 8, 512, 512, 3, 1),    # ResNet layer4 3x3
    # 384x128 inputs (every shipped YAML, BASELINE config 2): maps whose 128-pixel tiles span images with ragged rows
    (4, 96, 32, 32, 32, 3, 1),     # branch 0 at 96x32 (row shift, 4 rows per tile)
    (4, 48, 16, 64, 64, 3, 1),     # branch 1 at 48x16
    (4, 24, 8, 128, 128, 3, 1),    # branch 2 at 24x8: 2 images x 8 rows per tile
    (5, 24, 8, 128, 128, 3, 1),    # ... ragged: last tile holds one image
    (8, 12, 4, 256, 256, 3, 1),    # branch 3 at 12x4: 8 images x 4 rows per tile
    (11, 12, 4, 256, 256, 3, 1),   # ... ragged
    (4, 48, 16, 64, 128, 3, 2),    # transition stride-2 48x16 -> 24x8
    (4, 24, 8, 128, 256, 3, 2),    # transition stride-2 24x8 -> 12x4
    (4, 12, 4, 256, 128, 1, 1),    # fuse 1x1 at 12x4
    (2, 24, 8, 1024, 2048, 1, 1),  # ResNet-50 layer4 expand at 384x128
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=[str(s) for s in CONV_SHAPES])
def test_conv_tc_matches_fp32_reference(shape):
    n, h, w, cin, cout, k, s = shape
    g = torch.Generator(device="cpu").manual_seed(hash(shape) % 1000)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).to(DEV)
    wp = pack_w(wt)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    y = torch.full((n, ho, wo, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    tiles = _lib.query("bpb_conv_num_mtiles", n, ho, wo)
    stats = torch.full((tiles, 2, cout), float("nan"), device=DEV)
    _lib.call("bpb_conv_bf16", x.data_ptr(), wp.data_ptr(), y.data_ptr(), stats.data_ptr(), n, h, w, cin, cout, k, s, 0, S())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(cout, k, k, cin).permute(0, 3, 1, 2), None, s, pad)
    ref = ref.permute(0, 2, 3, 1)
    assert torch.isfinite(y.float()).all()
    err = rel_err(y.float(), ref)
    assert err < 1e-2, f"conv output rel err {err}"          # bf16 output rounding = 2^-9 relative + accumulation order
    # statistics are taken from the bf16-rounded outputs
    yb = y.float().reshape(-1, cout)
    s_ref, q_ref = yb.sum(0), (yb * yb).sum(0)
    st = stats.sum(0)
    assert rel_err(st[0], s_ref) < 1e-3 and rel_err(st[1], q_ref) < 1e-3
    # SIMT reference kernel agrees too (same operands)
    y2 = torch.empty_like(y)
    _lib.call("bpb_conv_ref_bf16", x.data_ptr(), wp.data_ptr(), y2.data_ptr(), n, h, w, cin, cout, k, s, S())
    torch.cuda.synchronize()
    assert rel_err(y2.float(), ref) < 1e-2


@pytest.mark.parametrize("shape", [(4, 64, 32, 32, 32, 3, 1), (4, 32, 16, 64, 64, 3, 1), (4, 16, 8, 128, 128, 3, 1),
                                   (2, 64, 32, 64, 64, 3, 1), (2, 64, 32, 256, 32, 3, 1)])
def test_conv_rowshift_mode_equals_single_tap_mode(shape):
    """3x3 stride-1 convs use the 'row shift' schedule (three vertical taps per TMA box); bn_hint bit 8 forces the
    one-tap-per-stage schedule.  Both must produce the same bf16 outputs (same products, same fp32 accumulation
    order up to tap grouping)."""
    n, h, w, cin, cout, k, s = shape
    g = torch.Generator().manual_seed(17)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(DEV)
    wp = pack_w((torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * 9)) ** 0.5)).to(DEV)
    ya = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    yb = torch.empty_like(ya)
    _lib.call("bpb_conv_bf16", x.data_ptr(), wp.data_ptr(), ya.data_ptr(), None, n, h, w, cin, cout, k, s, 0, S())
    _lib.call("bpb_conv_bf16", x.data_ptr(), wp.data_ptr(), yb.data_ptr(), None, n, h, w, cin, cout, k, s, 0x100, S())
    torch.cuda.synchronize()
    assert rel_err(ya.float(), yb.float()) < 4e-3          # one bf16 ulp where the fp32 sums round differently


def test_conv_fused_bn_matches_two_pass_statistics():
    """bpb_conv_bn_bf16 (last-CTA finalize) == conv + per-tile partials + bpb_bn_finalize, incl. running stats, and the
    workspace is left zeroed so the same buffer can be reused by the next launch."""
    n, h, w, cin, cout, k, s = 8, 32, 16, 64, 64, 3, 1
    g = torch.Generator().manual_seed(23)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(DEV)
    wp = pack_w(torch.randn(cout, cin, k, k, generator=g) * 0.05).to(DEV)
    gamma, beta = (torch.rand(cout, generator=g) + 0.5).to(DEV), torch.randn(cout, generator=g).to(DEV)
    y1, y2 = (torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16) for _ in range(2))
    ws = torch.zeros(_lib.query("bpb_conv_bn_workspace_bytes", cout), dtype=torch.uint8, device=DEV)
    rm1, rv1, rm2, rv2 = (torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV),
                          torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV))
    sc1, sh1, sc2, sh2 = (torch.empty(cout, device=DEV) for _ in range(4))
    for _ in range(2):      # twice: the second launch must find a clean workspace
        rm1.zero_(); rv1.fill_(1.0)
        _lib.call("bpb_conv_bn_bf16", x.data_ptr(), wp.data_ptr(), y1.data_ptr(), None, n, h, w, cin, cout, k, s, 0,
                  ws.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm1.data_ptr(), rv1.data_ptr(), 0,
                  sc1.data_ptr(), sh1.data_ptr(), None, S())
    tiles = _lib.query("bpb_conv_num_mtiles", n, h, w)
    stats = torch.empty(tiles, 2, cout, device=DEV)
    _lib.call("bpb_conv_bf16", x.data_ptr(), wp.data_ptr(), y2.data_ptr(), stats.data_ptr(), n, h, w, cin, cout, k, s, 0, S())
    _lib.call("bpb_bn_finalize", stats.data_ptr(), tiles, cout, 0, n * h * w, gamma.data_ptr(), beta.data_ptr(), 1e-5,
              0.1, rm2.data_ptr(), rv2.data_ptr(), sc2.data_ptr(), sh2.data_ptr(), None, S())
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    assert rel_err(sc1, sc2) < 1e-5 and rel_err(sh1, sh2) < 1e-5 and rel_err(rm1, rm2) < 1e-5 and rel_err(rv1, rv2) < 1e-5
    assert int(ws.count_nonzero()) == 0


def test_conv_bn_hint_variants():
    n, h, w, cin, cout, k, s = 4, 32, 16, 64, 128, 3, 1
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(DEV)
    wt = (torch.randn(cout, cin, k, k, generator=g) * 0.05).to(DEV)
    wp = pack_w(wt)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wp.float().view(cout, k, k, cin).permute(0, 3, 1, 2), None, s, 1)
    ref = ref.permute(0, 2, 3, 1)
    for bn in (32, 64, 128):
        y = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
        _lib.call("bpb_conv_bf16", x.data_ptr(), wp.data_ptr(), y.data_ptr(), None, n, h, w, cin, cout, k, s, bn, S())
        torch.cuda.synchronize()
        assert rel_err(y.float(), ref) < 1e-2, bn


@pytest.mark.parametrize("m,n,k,ldo", [(256, 512, 1920, 512), (1280, 512, 1920, 512), (256, 751, 512, 752),
                                        (8, 751, 512, 752), (40, 512, 2048, 512)])
def test_linear_tf32(m, n, k, ldo):
    g = torch.Generator().manual_seed(m + n)
    a = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) * 0.02).to(DEV)
    b = torch.randn(n, generator=g).to(DEV)
    out = torch.full((m, ldo), float("nan"), device=DEV)
    _lib.call("bpb_linear_tf32", a.data_ptr(), k, w.data_ptr(), b.data_ptr(), out.data_ptr(), m, n, k, ldo, 0, 0, S())
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().t() + b.double()).float()
    err = rel_err(out[:, :n], ref)
    assert err < 1e-3, err          # tf32 inputs (10-bit mantissa), fp32 accumulation


def test_linear_tf32_part_concat_mode():
    kparts, m, d, n = 5, 64, 512, 751
    g = torch.Generator().manual_seed(9)
    a = torch.randn(kparts, m, d, generator=g).to(DEV)
    w = (torch.randn(n, kparts * d, generator=g) * 0.02).to(DEV)
    out = torch.empty(m, 752, device=DEV)
    _lib.call("bpb_linear_tf32", a.data_ptr(), d, w.data_ptr(), None, out.data_ptr(), m, n, kparts * d, 752, 1, kparts, S())
    torch.cuda.synchronize()
    cat = a.permute(1, 0, 2).reshape(m, kparts * d)
    ref = (cat.double() @ w.double().t()).float()
    assert rel_err(out[:, :n], ref) < 1e-3


def test_bn_pipeline_and_fuse_and_upsample():
    n, h, w, c = 4, 32, 16, 64
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(n, h, w, c, generator=g) * 2 + 0.5).to(torch.bfloat16).to(DEV)
    tiles = _lib.query("bpb_colstats_tiles", n * h * w)
    part = torch.empty(tiles, 2, c, device=DEV)
    _lib.call("bpb_colstats_bf16", x.data_ptr(), part.data_ptr(), n * h * w, c, S())
    gamma = torch.rand(c, generator=g).to(DEV) + 0.5
    beta = torch.randn(c, generator=g).to(DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    scale, shift = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    _lib.call("bpb_bn_finalize", part.data_ptr(), tiles, c, 0, n * h * w, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1,
              rm.data_ptr(), rv.data_ptr(), scale.data_ptr(), shift.data_ptr(), None, S())
    res = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).to(DEV)
    y = torch.empty_like(x)
    _lib.call("bpb_bn_apply", x.data_ptr(), scale.data_ptr(), shift.data_ptr(), res.data_ptr(), None, None,
              y.data_ptr(), n * h * w, c, 1, S())
    torch.cuda.synchronize()
    xf = x.float().permute(0, 3, 1, 2)
    rm2, rv2 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    ref = F.relu(F.batch_norm(xf, rm2, rv2, gamma, beta, True, 0.1, 1e-5) + res.float().permute(0, 3, 1, 2))
    assert rel_err(y.float().permute(0, 3, 1, 2), ref) < 1e-2
    assert rel_err(rm, rm2) < 1e-4 and rel_err(rv, rv2) < 1e-4           # running statistics side effect
    # fuse: y = relu(a + up2(b*s+t) + up4(c))
    a = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).to(DEV)
    b = torch.randn(n, h // 2, w // 2, c, generator=g).to(torch.bfloat16).to(DEV)
    cc = torch.randn(n, h // 4, w // 4, c, generator=g).to(torch.bfloat16).to(DEV)
    srcs = (ctypes.c_void_p * 3)(a.data_ptr(), b.data_ptr(), cc.data_ptr())
    scs = (ctypes.c_void_p * 3)(None, scale.data_ptr(), None)
    shs = (ctypes.c_void_p * 3)(None, shift.data_ptr(), None)
    ups = (ctypes.c_int * 3)(0, 1, 2)
    out = torch.empty_like(a)
    _lib.call("bpb_fuse_sum", ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(scs, ctypes.c_void_p),
              ctypes.cast(shs, ctypes.c_void_p), ctypes.cast(ups, ctypes.c_void_p), 3, out.data_ptr(), n, h, w, c, 1, S())
    torch.cuda.synchronize()
    bb = (b.float() * scale + shift).permute(0, 3, 1, 2)
    ref = F.relu(a.float().permute(0, 3, 1, 2) + F.interpolate(bb, scale_factor=2, mode="nearest") +
                 F.interpolate(cc.float().permute(0, 3, 1, 2), scale_factor=4, mode="nearest"))
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 1e-2
    # bilinear(align_corners) upsample + concat
    srcs = (ctypes.c_void_p * 3)(a.data_ptr(), b.data_ptr(), cc.data_ptr())
    chans = (ctypes.c_int * 3)(c, c, c)
    hs = (ctypes.c_int * 3)(h, h // 2, h // 4)
    ws = (ctypes.c_int * 3)(w, w // 2, w // 4)
    cat = torch.empty(n, h, w, 3 * c, device=DEV, dtype=torch.bfloat16)
    _lib.call("bpb_upsample_concat", ctypes.cast(srcs, ctypes.c_void_p), ctypes.cast(chans, ctypes.c_void_p),
              ctypes.cast(hs, ctypes.c_void_p), ctypes.cast(ws, ctypes.c_void_p), 3, cat.data_ptr(), n, h, w, S())
    torch.cuda.synchronize()
    up = lambda t: F.interpolate(t.float().permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=True)  # noqa: E731
    ref = torch.cat([a.float().permute(0, 3, 1, 2), up(b), up(cc)], 1)
    assert rel_err(cat.float().permute(0, 3, 1, 2), ref) < 1e-2


def test_stem_and_maxpool():
    n, h, w = 3, 64, 32
    g = torch.Generator().manual_seed(8)
    img = torch.randn(n, 3, h, w, generator=g).to(DEV)
    for k in (3, 7):
        wt = (torch.randn(64, 3, k, k, generator=g) * 0.1).to(DEV)
        ho, wo = (h + 2 * (k // 2) - k) // 2 + 1, (w + 2 * (k // 2) - k) // 2 + 1
        y = torch.empty(n, ho, wo, 64, device=DEV, dtype=torch.bfloat16)
        rows = _lib.query("bpb_stem_conv_tiles", n, ho, wo)
        part = torch.full((rows, 2, 64), float("nan"), device=DEV)
        _lib.call("bpb_stem_conv", img.data_ptr(), wt.data_ptr(), y.data_ptr(), part.data_ptr(), n, h, w, 64, k, 2, S())
        torch.cuda.synchronize()
        ref = F.conv2d(img.to(torch.bfloat16).float(), wt, None, 2, k // 2)
        assert rel_err(y.float().permute(0, 3, 1, 2), ref) < 1e-2
        yb = y.float().reshape(-1, 64)
        st = part.sum(0)
        assert rel_err(st[0], yb.sum(0)) < 1e-3 and rel_err(st[1], (yb * yb).sum(0)) < 1e-3
    x = torch.randn(n, 33, 17, 64, generator=g).to(torch.bfloat16).to(DEV)
    for pad, ceil in ((1, 0), (0, 1)):
        ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, pad, ceil_mode=bool(ceil))
        y = torch.empty(n, ref.shape[2], ref.shape[3], 64, device=DEV, dtype=torch.bfloat16)
        _lib.call("bpb_maxpool3x3s2", x.data_ptr(), y.data_ptr(), n, 33, 17, 64, pad, ceil, S())
        torch.cuda.synchronize()
        assert torch.equal(y.float().permute(0, 3, 1, 2), ref)


def test_ibn_kernels():
    n, hw, c, half = 3, 128, 64, 32
    g = torch.Generator().manual_seed(4)
    x = torch.randn(n, hw, c, generator=g).to(torch.bfloat16).to(DEV)
    mr = torch.empty(n, half, 2, device=DEV)
    _lib.call("bpb_instnorm_stats", x.data_ptr(), mr.data_ptr(), n, hw, c, 0, half, 1e-5, S())
    ig, ib = torch.rand(half, generator=g).to(DEV) + 0.5, torch.randn(half, generator=g).to(DEV)
    sc, sh = torch.rand(c, generator=g).to(DEV) + 0.5, torch.randn(c, generator=g).to(DEV)
    y = torch.empty_like(x)
    _lib.call("bpb_ibn_apply", x.data_ptr(), mr.data_ptr(), ig.data_ptr(), ib.data_ptr(), sc.data_ptr(), sh.data_ptr(),
              y.data_ptr(), n, hw, c, half, 1, S())
    torch.cuda.synchronize()
    xf = x.float().permute(0, 2, 1).reshape(n, c, hw, 1)
    a = F.instance_norm(xf[:, :half], None, None, ig, ib, True, 0.0, 1e-5)
    b = xf[:, half:] * sc[half:].view(1, -1, 1, 1) + sh[half:].view(1, -1, 1, 1)
    ref = F.relu(torch.cat([a, b], 1)).reshape(n, c, hw).permute(0, 2, 1)
    assert rel_err(y.float(), ref) < 1e-2
