"""Forward-pass planner: turns a BPBreID parameter tree + an input shape into a flat list of C-ABI launches
(libbpbreid_sm100.so) over statically allocated device buffers, so that one training / inference step is a
replayable sequence (eager loop or CUDA graph).  Python only orchestrates: every number is produced by the
hand-written sm_100a kernels.

Reference call stack replaced: BPBreID.forward (torchreid/models/bpbreid.py:116-259) ->
HighResolutionNet.forward (torchreid/models/hrnet.py:532-576) / ResNet.featuremaps (torchreid/models/resnet.py:342-352)
/ fast-reid ResNet.forward (torchreid/models/resnet_fastreid.py:296-345).

Data layout: activations NHWC bf16; conv weights OHWI bf16 (packed once per plan); BN runs in two phases
(conv epilogue emits per-tile partial sums -> bn_finalize -> scale/shift consumed by the next elementwise pass);
pooled embeddings are group-major fp32 [K+3, N, C] = {global, foreground, background, part_0..K-1}.
"""
from __future__ import annotations

import ctypes
import os
from contextlib import contextmanager

import torch
from torch import nn

from . import _lib
from . import params as P

BN_EPS = 1e-5


class Act:
    """NHWC bf16 activation."""
    __slots__ = ("t", "n", "h", "w", "c", "buf")

    def __init__(self, t, n, h, w, c, buf=None):
        self.t, self.n, self.h, self.w, self.c, self.buf = t, n, h, w, c, buf

    @property
    def m(self):
        return self.n * self.h * self.w


class Lazy:
    """Raw conv output + the BN scale/shift that still has to be applied to it.  In gradient mode it also remembers
    how it was produced (conv module, its input, the BN module, saved batch mean / invstd) for the backward pass."""
    __slots__ = ("raw", "scale", "shift", "conv", "x_in", "bn", "save", "kind")

    def __init__(self, raw, scale, shift, conv=None, x_in=None, bn=None, save=None, kind="tc"):
        self.raw, self.scale, self.shift = raw, scale, shift
        self.conv, self.x_in, self.bn, self.save, self.kind = conv, x_in, bn, save, kind


def _tc_tileable(cin, cout, ho, wo):
    """Shapes the tcgen05 tiler accepts: 128 output pixels = (images) x (rows) x (full output width); the rows per
    tile are the largest divisor of ho that keeps the tile at 128 pixels (tc_gemm.cu: conv_tiling)."""
    return not (cin % 32 or cout % 32 or wo > 128 or 128 % wo)


class ForwardPlan:
    def __init__(self, model, n, h, w, training, device, conv_impl="tc", use_graph=False, debug=False, grad=False,
                 flat=None, ext_shape=None):
        self.model = model
        # external part masks [N, K+1, Hm, Wm] (bpbreid.py:116-175): mode bit 0 = they are the attention
        # (learnable_attention_enabled False), bit 1 / 2 = test_use_target_segmentation 'hard' / 'soft' (eval only)
        cfg_ = model.model_cfg
        self.ext_mode = 0 if cfg_.learnable_attention_enabled else 1
        if not training and cfg_.test_use_target_segmentation == "hard":
            self.ext_mode |= 2
        if not training and cfg_.test_use_target_segmentation == "soft":
            self.ext_mode |= 4
        self.ext_shape = ext_shape
        if self.ext_mode and ext_shape is None:
            raise AssertionError("external_parts_masks is required by this configuration (bpbreid.py:151,162,171)")
        # gradient mode (bpbreid_b200.train): every activation stays alive for the backward pass, BatchNorm layers
        # save their batch mean / invstd, conv weights come from the flat parameter store's bf16 operand copies
        self.grad = bool(grad)
        self.flat = flat
        self.tape = []
        self.fuse_eval = getattr(model, "fuse_eval_epilogue", True)
        self.n, self.h, self.w = n, h, w
        self.training = bool(training)
        self.device = device
        self.conv_impl = conv_impl
        self.ops = []                 # (ctypes fn, args, name, lane) ; stream appended at run time; fn None = fork/join
        self._lane = 0                # stream lane ops are being recorded on (0 = the caller's stream)
        self._buf_lane = {}           # id(buffer) -> lane that owns it (buffers are only reused inside their lane)
        self._side_streams = []
        self.setup_ops = []           # run once at build / after a weight refresh (eval-mode BN folds)
        self.keepalive = []           # ctypes arrays referenced by ops
        self.persistent = []          # tensors that must outlive the plan build
        self._free = {}               # pooled activation buffers by byte size
        self._scopes = [[]]
        self.launches = 0
        self.graph = None
        self.use_graph = use_graph
        self.multi_stream = os.environ.get("BPB_SINGLE_STREAM", "0") != "1"
        self.lib = _lib.load()
        # debug: keep every intermediate activation alive (no buffer reuse) and record (block, inputs, outputs) so
        # tests can check each block of the real network against the oracle on the block's actual input
        self.debug = debug
        self.trace = []
        self._names = {id(mod): name for name, mod in model.named_modules()} if debug else {}
        self.cfg = model.model_cfg
        self.K = model.parts_num
        self.images = torch.zeros(n, 3, h, w, device=device, dtype=torch.float32)
        self._packed = {}
        self._need_setup = True
        with torch.no_grad():
            self._build()

    # ------------------------------------------------------------------ buffers
    def _alloc(self, shape, dtype=torch.bfloat16):
        numel = 1
        for s in shape:
            numel *= s
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        lst = self._free.get((self._lane, nbytes))
        if lst:
            buf = lst.pop()
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._buf_lane[id(buf)] = self._lane
        self._scopes[-1].append(buf)
        return buf.view(dtype).view(*shape), buf

    def _act(self, n, h, w, c):
        t, buf = self._alloc((n, h, w, c))
        return Act(t, n, h, w, c, buf)

    def _small(self, shape, dtype=torch.float32, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self.persistent.append(t)
        return t

    @contextmanager
    def scope(self):
        """Activations allocated inside are returned to the pool on exit unless kept."""
        self._scopes.append([])
        keep = []
        try:
            yield keep
        finally:
            mine = self._scopes.pop()
            kept_ids = {id(a.buf) for a in keep}
            for buf in mine:
                if self.debug or self.grad or id(buf) in kept_ids:
                    self._scopes[-1].append(buf)
                else:
                    self._free.setdefault((self._buf_lane[id(buf)], buf.numel()), []).append(buf)

    def _trace(self, kind, mods, ins, outs, **extra):
        if self.debug:
            names = [self._names.get(id(m), "?") for m in (mods if isinstance(mods, (list, tuple)) else [mods])]
            self.trace.append(dict(kind=kind, names=names, ins=list(ins), outs=list(outs), **extra))

    # ------------------------------------------------------------------ op recording
    def _op(self, name, *args):
        self.ops.append((getattr(self.lib, name), args, name, self._lane))
        self.launches += 1

    # ---- stream lanes: independent sub-graphs (HRNet branches, fuse outputs) are recorded on side streams so that
    # ---- the small low-resolution layers overlap the big ones; buffers never migrate between lanes while live
    @contextmanager
    def lane(self, idx):
        prev, self._lane = self._lane, (idx if self.multi_stream else 0)
        try:
            yield
        finally:
            self._lane = prev

    def fork(self, lanes):
        """side lanes start after everything recorded so far on lane 0"""
        if self.multi_stream:
            lanes = tuple(l for l in lanes if l != 0)
            self.ops.append((None, lanes, "fork", 0))
            self._open_lanes = lanes
            if self.grad:
                self.tape.append(("fork", lanes))

    def join(self):
        """lane 0 waits for every side lane forked since the last join"""
        if self.multi_stream:
            self.ops.append((None, (), "join", 0))
            if self.grad:
                self.tape.append(("join", getattr(self, "_open_lanes", ())))

    def _setup_op(self, name, *args):
        self.setup_ops.append((getattr(self.lib, name), args, name))

    def _ptr_array(self, tensors):
        arr = (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
        self.keepalive.append(arr)
        return ctypes.cast(arr, ctypes.c_void_p)

    def _int_array(self, vals):
        arr = (ctypes.c_int * len(vals))(*vals)
        self.keepalive.append(arr)
        return ctypes.cast(arr, ctypes.c_void_p)

    def _float_array(self, vals):
        arr = (ctypes.c_float * len(vals))(*vals)
        self.keepalive.append(arr)
        return ctypes.cast(arr, ctypes.c_void_p)

    # ------------------------------------------------------------------ weights
    def pack_conv(self, conv: nn.Conv2d):
        key = id(conv)
        if self.flat is not None:
            return self.flat.forward_operand(conv)
        if key not in self._packed:
            w = conv.weight.detach()
            self._packed[key] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.bfloat16).contiguous()
        return self._packed[key]

    def refresh_weights(self):
        """Re-pack the bf16 OHWI copies after the fp32 master weights changed."""
        for mod in self.model.modules():
            if isinstance(mod, nn.Conv2d) and id(mod) in self._packed:
                w = mod.weight.detach()
                self._packed[id(mod)].copy_(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
        self._need_setup = True      # eval-mode BN folds depend on the (possibly reloaded) running statistics

    # ------------------------------------------------------------------ building blocks
    def _bn_params(self, bn, c_lo=0):
        return bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum if bn.momentum is not None else 0.1

    def bn_scale_shift(self, stats, tiles, c, count, bn, c_lo=0):
        """partials -> (scale, shift) for nn.BatchNorm2d `bn` (train: batch stats, eval: running stats)."""
        scale, shift = self._small((c,)), self._small((c,))
        g, b, rm, rv, mom = self._bn_params(bn)
        self._last_save = self._small((2, c)) if (self.grad and self.training) else None
        if self.training:
            self._op("bpb_bn_finalize", stats.data_ptr(), tiles, c, c_lo, count, g.data_ptr(), b.data_ptr(),
                     BN_EPS, float(mom), rm.data_ptr(), rv.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                     _lib.ptr(self._last_save))
        else:
            # inference: running statistics are constants -> folded once, outside the replayed step
            assert c_lo == 0
            self._setup_op("bpb_bn_eval_fold", g.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(), BN_EPS, c,
                           scale.data_ptr(), shift.data_ptr())
        return scale, shift

    def conv_raw(self, x: Act, conv: nn.Conv2d, need_stats=True):
        k, s = conv.kernel_size[0], conv.stride[0]
        cout = conv.out_channels
        pad = k // 2
        ho = (x.h + 2 * pad - k) // s + 1
        wo = (x.w + 2 * pad - k) // s + 1
        y = self._act(x.n, ho, wo, cout)
        wp = self.pack_conv(conv)
        tc_ok = self.conv_impl == "tc" and _tc_tileable(x.c, cout, ho, wo)
        stats, tiles = None, 0
        if tc_ok:
            tiles = _lib.query("bpb_conv_num_mtiles", x.n, ho, wo)
            if need_stats and self.training:
                stats = self._small((tiles, 2, cout))
            self._op("bpb_conv_bf16", x.t.data_ptr(), wp.data_ptr(), y.t.data_ptr(), _lib.ptr(stats), x.n, x.h, x.w,
                     x.c, cout, k, s, 0)
        else:
            self._op("bpb_conv_ref_bf16", x.t.data_ptr(), wp.data_ptr(), y.t.data_ptr(), x.n, x.h, x.w, x.c, cout, k, s)
            if need_stats and self.training:
                stats, tiles = self.colstats(y)
        return y, stats, tiles

    def colstats(self, y: Act):
        tiles = _lib.query("bpb_colstats_tiles", y.m)
        stats = self._small((tiles, 2, y.c))
        self._op("bpb_colstats_bf16", y.t.data_ptr(), stats.data_ptr(), y.m, y.c)
        return stats, tiles

    def conv_bn(self, x: Act, conv, bn, c_lo=0) -> Lazy:
        """conv + BatchNorm statistics.  Training + tensor-core path: ONE launch (the conv's last CTA finalizes the
        batch statistics into scale/shift and updates the running stats); otherwise conv, then bn_finalize / fold.
        Inference + tensor-core path: nothing is launched yet -- the returned Lazy is PENDING (raw is None) and the
        consumer (`apply`, `fuse`) launches the convolution with BatchNorm / residual / ReLU in its epilogue."""
        k, s = conv.kernel_size[0], conv.stride[0]
        cout, pad = conv.out_channels, k // 2
        ho, wo = (x.h + 2 * pad - k) // s + 1, (x.w + 2 * pad - k) // s + 1
        if (not self.training and not self.debug and c_lo == 0 and self.conv_impl == "tc" and self.fuse_eval
                and _tc_tileable(x.c, cout, ho, wo)):
            scale, shift = self._small((cout,)), self._small((cout,))
            self._setup_op("bpb_bn_eval_fold", bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                           bn.running_var.data_ptr(), BN_EPS, cout, scale.data_ptr(), shift.data_ptr())
            return Lazy(None, scale, shift, conv, x, bn, None, "pending")
        if self.training and self.conv_impl == "tc" and _tc_tileable(x.c, cout, ho, wo):
            y = self._act(x.n, ho, wo, cout)
            wp = self.pack_conv(conv)
            ws = torch.zeros(_lib.query("bpb_conv_bn_workspace_bytes", cout), dtype=torch.uint8, device=self.device)
            self.persistent.append(ws)
            scale, shift = self._small((cout,)), self._small((cout,))
            save = self._small((2, cout)) if self.grad else None
            mom = bn.momentum if bn.momentum is not None else 0.1
            self._op("bpb_conv_bn_bf16", x.t.data_ptr(), wp.data_ptr(), y.t.data_ptr(), None, x.n, x.h, x.w, x.c, cout,
                     k, s, 0, ws.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), BN_EPS, float(mom),
                     bn.running_mean.data_ptr(), bn.running_var.data_ptr(), c_lo, scale.data_ptr(), shift.data_ptr(),
                     _lib.ptr(save))
            return Lazy(y, scale, shift, conv, x, bn, save, "tc")
        y, stats, tiles = self.conv_raw(x, conv)
        if self.grad and self.training:
            raise NotImplementedError(f"training backward needs the tensor-core conv path (shape {x.h}x{x.w}x{x.c}->{cout})")
        if c_lo == 0:
            scale, shift = self.bn_scale_shift(stats, tiles, y.c, y.m, bn)
        else:
            scale, shift = self._small((y.c,)), self._small((y.c,))
            if self.training:
                mom = bn.momentum if bn.momentum is not None else 0.1
                self._op("bpb_bn_finalize", stats.data_ptr(), tiles, y.c, c_lo, y.m, bn.weight.data_ptr(),
                         bn.bias.data_ptr(), BN_EPS, float(mom), bn.running_mean.data_ptr(),
                         bn.running_var.data_ptr(), scale.data_ptr(), shift.data_ptr(), None)
            else:
                self._setup_op("bpb_bn_eval_fold", bn.weight.data_ptr(), bn.bias.data_ptr(),
                               bn.running_mean.data_ptr(), bn.running_var.data_ptr(), BN_EPS, y.c - c_lo,
                               scale[c_lo:].data_ptr(), shift[c_lo:].data_ptr())
        return Lazy(y, scale, shift)

    def fused_conv(self, lz: Lazy, relu, res: Act | None = None) -> Act:
        """launch a pending conv with BatchNorm (+ residual) (+ ReLU) in its epilogue (inference)"""
        x, conv = lz.x_in, lz.conv
        k, s = conv.kernel_size[0], conv.stride[0]
        cout, pad = conv.out_channels, k // 2
        ho, wo = (x.h + 2 * pad - k) // s + 1, (x.w + 2 * pad - k) // s + 1
        y = self._act(x.n, ho, wo, cout)
        self._op("bpb_conv_bn_act_bf16", x.t.data_ptr(), self.pack_conv(conv).data_ptr(), y.t.data_ptr(), x.n, x.h, x.w,
                 x.c, cout, k, s, lz.scale.data_ptr(), lz.shift.data_ptr(), None if res is None else res.t.data_ptr(),
                 1 if relu else 0)
        return y

    def launch_raw(self, lz: Lazy) -> Lazy:
        """a pending conv whose consumer combines it in fp32 with other terms (projection skip, fuse-layer sum): run the
        plain conv now (raw bf16 output), scale / shift stay with the consumer exactly like in the two-pass path"""
        if lz.kind != "pending":
            return lz
        y, _, _ = self.conv_raw(lz.x_in, lz.conv, need_stats=False)
        return Lazy(y, lz.scale, lz.shift, lz.conv, lz.x_in, lz.bn, None, "tc")

    def apply(self, lz: Lazy, relu, res=None) -> Act:
        """y = act(bn(raw) + res) ; res may be an Act (identity skip) or a Lazy (projection skip)."""
        if isinstance(res, Lazy):                              # projection skip: fp32 combination in bn_apply
            res = self.launch_raw(res)
            lz = self.launch_raw(lz)
        if lz.kind == "pending":
            return self.fused_conv(lz, relu, res)
        r = lz.raw
        y = self._act(r.n, r.h, r.w, r.c)
        if res is None:
            rp, rs, rb = None, None, None
        elif isinstance(res, Lazy):
            rp, rs, rb = res.raw.t.data_ptr(), res.scale.data_ptr(), res.shift.data_ptr()
        else:
            rp, rs, rb = res.t.data_ptr(), None, None
        self._op("bpb_bn_apply", r.t.data_ptr(), lz.scale.data_ptr(), lz.shift.data_ptr(), rp, rs, rb, y.t.data_ptr(),
                 r.m, r.c, 1 if relu else 0)
        return y

    def basic_block(self, x: Act, blk) -> Act:
        with self.scope() as keep:
            a = self.apply(self.conv_bn(x, blk.conv1, blk.bn1), True)
            y = self.apply(self.conv_bn(a, blk.conv2, blk.bn2), True, res=x)
            keep.append(y)
        self._trace("basic", blk, [x], [y])
        return y

    def bottleneck(self, x: Act, blk) -> Act:
        with self.scope() as keep:
            if isinstance(blk.bn1, P.IBNPair):
                a = self.ibn_relu(x, blk.conv1, blk.bn1)
            else:
                a = self.apply(self.conv_bn(x, blk.conv1, blk.bn1), True)
            b = self.apply(self.conv_bn(a, blk.conv2, blk.bn2), True)
            l3 = self.conv_bn(b, blk.conv3, blk.bn3)
            if blk.downsample is not None:
                res = self.conv_bn(x, blk.downsample[0], blk.downsample[1])
            else:
                res = x
            y = self.apply(l3, True, res=res)
            keep.append(y)
        self._trace("bottleneck", blk, [x], [y], stride=blk.stride)
        return y

    def ibn_relu(self, x: Act, conv, ibn) -> Act:
        """conv -> IBN (InstanceNorm first half | BatchNorm second half) -> ReLU (resnet_fastreid.py:108-122,224-226)."""
        half = ibn.half
        lz = self.conv_bn(x, conv, ibn.BN, c_lo=half)
        raw, c = lz.raw, lz.raw.c
        mr = self._small((raw.n, half, 2))
        self._op("bpb_instnorm_stats", raw.t.data_ptr(), mr.data_ptr(), raw.n, raw.h * raw.w, c, 0, half, BN_EPS)
        y = self._act(raw.n, raw.h, raw.w, c)
        self._op("bpb_ibn_apply", raw.t.data_ptr(), mr.data_ptr(), ibn.IN.weight.data_ptr(), ibn.IN.bias.data_ptr(),
                 lz.scale.data_ptr(), lz.shift.data_ptr(), y.t.data_ptr(), raw.n, raw.h * raw.w, c, half, 1)
        return y

    def fuse(self, terms, out_like: Act, relu=True) -> Act:
        """terms: list of (Act | Lazy, log2 upsample factor)"""
        terms = [(self.launch_raw(t), up) if isinstance(t, Lazy) else (t, up) for t, up in terms]
        y = self._act(out_like.n, out_like.h, out_like.w, out_like.c)
        srcs, scs, shs, ups = [], [], [], []
        for t, up in terms:
            if isinstance(t, Lazy):
                srcs.append(t.raw.t); scs.append(t.scale); shs.append(t.shift)
            else:
                srcs.append(t.t); scs.append(None); shs.append(None)
            ups.append(up)
        self._op("bpb_fuse_sum", self._ptr_array(srcs), self._ptr_array(scs), self._ptr_array(shs),
                 self._int_array(ups), len(terms), y.t.data_ptr(), y.n, y.h, y.w, y.c, 1 if relu else 0)
        return y

    def stem(self, conv, bn, ksize, stride) -> Act:
        n, h, w = self.n, self.h, self.w
        pad = ksize // 2
        ho, wo = (h + 2 * pad - ksize) // stride + 1, (w + 2 * pad - ksize) // stride + 1
        raw = self._act(n, ho, wo, conv.out_channels)
        wt = conv.weight.detach()
        stats, tiles = None, 0
        if self.training:
            tiles = _lib.query("bpb_stem_conv_tiles", n, ho, wo)
            stats = self._small((tiles, 2, conv.out_channels))
        self._op("bpb_stem_conv", self.images.data_ptr(), wt.data_ptr(), raw.t.data_ptr(), _lib.ptr(stats), n, h, w,
                 conv.out_channels, ksize, stride)
        scale, shift = self.bn_scale_shift(stats, tiles, raw.c, raw.m, bn)
        lz = Lazy(raw, scale, shift, conv, None, bn, self._last_save, "stem")
        self._stem_geom = (ksize, stride)
        y = self.apply(lz, True)
        self._trace("stem", [conv, bn], [], [y], ksize=ksize, stride=stride)
        return y

    def conv_bn_relu(self, x: Act, conv, bn) -> Act:
        y = self.apply(self.conv_bn(x, conv, bn), True)
        self._trace("conv_bn_relu", [conv, bn], [x], [y], stride=conv.stride[0], ksize=conv.kernel_size[0])
        return y

    # ------------------------------------------------------------------ backbones
    def hr_module(self, xs, mod):
        nb = len(xs)
        with self.scope() as keep:
            # branch b runs on stream lane b (independent until the fuse layer, hrnet.py:259-279)
            ys = []
            self.fork(range(1, nb))
            for b in range(nb):
                with self.lane(b):
                    t = xs[b]
                    for blk in mod.branches[b]:
                        t = self.basic_block(t, blk)
                    ys.append(t)
            self.join()
            # fuse output i (needs every branch) on lane i
            outs = []
            self.fork(range(1, nb))
            for i in range(nb):
                with self.lane(i):
                    terms = []
                    for j in range(nb):
                        if j == i:
                            terms.append((ys[j], 0))
                        elif j > i:
                            f = mod.fuse_layers[i][j]
                            terms.append((self.conv_bn(ys[j], f[0], f[1]), j - i))
                        else:
                            t = ys[j]
                            chain = mod.fuse_layers[i][j]
                            for k, st in enumerate(chain):
                                lz = self.conv_bn(t, st[0], st[1])
                                if k == len(chain) - 1:
                                    terms.append((lz, 0))
                                else:
                                    t = self.apply(lz, True)
                    outs.append(self.fuse(terms, ys[i]))
            self.join()
            keep.extend(outs)
        self._trace("fuse", mod, ys, outs)
        return outs

    def build_hrnet(self, bb):
        with self.scope() as keep:
            x = self.stem(bb.conv1, bb.bn1, 3, 2)
            x = self.conv_bn_relu(x, bb.conv2, bb.bn2)
            for blk in bb.layer1:
                x = self.bottleneck(x, blk)
            t = bb.transition1
            xs = [self.conv_bn_relu(x, t[0][0], t[0][1]), self.conv_bn_relu(x, t[1][0][0], t[1][0][1])]
            self.mark("stage2")
            for mod in bb.stage2:
                xs = self.hr_module(xs, mod)
            t = bb.transition2[2][0]
            xs = xs + [self.conv_bn_relu(xs[-1], t[0], t[1])]
            self.mark("stage3")
            for mod in bb.stage3:
                xs = self.hr_module(xs, mod)
            t = bb.transition3[3][0]
            xs = xs + [self.conv_bn_relu(xs[-1], t[0], t[1])]
            self.mark("stage4")
            for mod in bb.stage4:
                xs = self.hr_module(xs, mod)
            br = []
            self.fork(range(1, 4))
            for i in range(4):
                with self.lane(i):
                    br.append(self.bottleneck(xs[i], bb.incre_modules[i][0]))
            self.join()
            keep.extend(br)
        return br

    def mark(self, name):
        """position marker (gradient mode: the backward of everything recorded AFTER this point is complete when the
        reversed tape reaches it -> a gradient bucket can be all-reduced while the earlier layers' backward runs)"""

    def hrnet_concat(self, br) -> Act:
        """bilinear(align_corners) upsample + concat (hrnet.py:568-573): only needed when the caller wants the
        `spatial_features` element of the 6-tuple or the map-based head."""
        feats = self._act(self.n, br[0].h, br[0].w, sum(b.c for b in br))
        self._op("bpb_upsample_concat", self._ptr_array([b.t for b in br]), self._int_array([b.c for b in br]),
                 self._int_array([b.h for b in br]), self._int_array([b.w for b in br]), 4, feats.t.data_ptr(),
                 self.n, feats.h, feats.w)
        self._trace("upcat", [], br, [feats])
        return feats

    def build_resnet(self, bb):
        with self.scope() as keep:
            x = self.stem(bb.conv1, bb.bn1, 7, 2)
            pad, ceil = (0, 1) if bb.fastreid else (1, 0)
            num_h, num_w = x.h + 2 * pad - 3, x.w + 2 * pad - 3
            ho = ((num_h + 1) // 2 if ceil else num_h // 2) + 1
            wo = ((num_w + 1) // 2 if ceil else num_w // 2) + 1
            if ceil and (ho - 1) * 2 >= x.h + pad:
                ho -= 1
            if ceil and (wo - 1) * 2 >= x.w + pad:
                wo -= 1
            y = self._act(x.n, ho, wo, x.c)
            self._op("bpb_maxpool3x3s2", x.t.data_ptr(), y.t.data_ptr(), x.n, x.h, x.w, x.c, pad, ceil)
            self._trace("maxpool", [], [x], [y], pad=pad, ceil=ceil)
            x = y
            for li in range(1, 5):
                for blk in getattr(bb, f"layer{li}"):
                    x = self.bottleneck(x, blk)
            keep.append(x)
        return x

    # ------------------------------------------------------------------ head
    def _ext_resized(self, h, w):
        """external masks -> feature-map resolution (recorded resize launch); None when the configuration ignores them"""
        if not self.ext_mode:
            return None
        n, kp1 = self.n, self.K + 1
        _, c_, hm, wm = self.ext_shape
        assert c_ == kp1, f"external_parts_masks must have K+1={kp1} channels, got {c_}"
        self.ext_in = self._small((n, kp1, hm, wm))
        self.ext_rs = self._small((n, kp1, h, w))
        self._op("bpb_resize_bilinear_f32", self.ext_in.data_ptr(), self.ext_rs.data_ptr(), n * kp1, hm, wm, h, w)
        return self.ext_rs

    def _alloc_head_outputs(self, h, w, c):
        n, kp1, J = self.n, self.K + 1, self.K + 3
        self.logits = self._small((n, kp1, h, w))
        self.probs = self._small((n, kp1, h, w))
        self.fg_mask = self._small((n, h, w))
        self.vis_bits = self._small((n,), torch.int32)
        self.vis_cont = self._small((n, kp1))
        self.vis_b = self._small((n, kp1 + 1), torch.uint8)
        self.vis_f = self._small((n, kp1 + 1))
        self.pooled = self._small((J, n, c))

    def build_head_front_map(self, feats: Act):
        """Pixel classifier / masks / pooling on a materialised feature map (ResNet backbones, or HRNet with
        lowres_head disabled)."""
        m, n = self.model, self.n
        kp1 = self.K + 1
        hw, c = feats.h * feats.w, feats.c
        pc = m.pixel_classifier
        learn = (self.ext_mode & 1) == 0
        wf = bf = None
        if learn:
            if self.training:
                stats, tiles = self.colstats(Act(feats.t, n, feats.h, feats.w, c))
            else:
                stats, tiles = None, 0
            scale, shift = self.bn_scale_shift(stats, tiles, c, n * hw, pc.bn)
            wf, bf = self._small((kp1, c)), self._small((kp1,))
            fold = self._op if self.training else self._setup_op
            fold("bpb_pixel_fold", pc.classifier.weight.data_ptr(), pc.classifier.bias.data_ptr(), scale.data_ptr(),
                 shift.data_ptr(), kp1, c, wf.data_ptr(), bf.data_ptr())
        ext = self._ext_resized(feats.h, feats.w)
        self._alloc_head_outputs(feats.h, feats.w, c)
        self._op("bpb_pixel_classify", feats.t.data_ptr(), _lib.ptr(wf), _lib.ptr(bf), n, hw, c, kp1,
                 self.logits.data_ptr(), self.probs.data_ptr(), self.fg_mask.data_ptr(), self.vis_bits.data_ptr(),
                 self.vis_cont.data_ptr(), _lib.ptr(ext), self.ext_mode)
        self._op("bpb_vis_unpack", self.vis_bits.data_ptr(), self.vis_cont.data_ptr(), n, kp1, self.vis_b.data_ptr(),
                 self.vis_f.data_ptr())
        self._op("bpb_part_pool", feats.t.data_ptr(), self.probs.data_ptr(), self.fg_mask.data_ptr(), n, hw, c, kp1,
                 self.pooled.data_ptr())

    def build_head_front_branches(self, br):
        """HRNet: the same head evaluated on the four un-concatenated branch maps (csrc/head_lowres.cu): bilinear
        upsampling, the 1x1 classifier and the pooling sums are linear, so nothing at 1920 channels x stride 4 is
        ever written.  BN2d batch statistics of the virtual upsampled map are accumulated on the fly."""
        m, n = self.model, self.n
        kp1, J = self.K + 1, self.K + 3
        h, w = br[0].h, br[0].w
        ctot = sum(b.c for b in br)
        offs = [sum(b.c for b in br[:i]) for i in range(len(br))]
        pc = m.pixel_classifier
        learn = (self.ext_mode & 1) == 0
        scale = shift = wf = bf = None
        lows = []
        if learn:
            scale, shift = self._small((ctot,)), self._small((ctot,))
            if self.training:
                partial = self._small((n, 2, ctot))
                self.fork(range(1, len(br)))
                for i, (b, off) in enumerate(zip(br, offs)):
                    with self.lane(i):
                        self._op("bpb_head_branch_stats", b.t.data_ptr(), n, b.h, b.w, b.c, h, w, partial.data_ptr(), off,
                                 ctot)
                self.join()
                bn = pc.bn
                mom = bn.momentum if bn.momentum is not None else 0.1
                self.pix_save = self._small((2, ctot)) if self.grad else None
                self._op("bpb_bn_finalize", partial.data_ptr(), n, ctot, 0, n * h * w, bn.weight.data_ptr(),
                         bn.bias.data_ptr(), BN_EPS, float(mom), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                         scale.data_ptr(), shift.data_ptr(), _lib.ptr(self.pix_save))
            else:
                self._setup_op("bpb_bn_eval_fold", pc.bn.weight.data_ptr(), pc.bn.bias.data_ptr(),
                               pc.bn.running_mean.data_ptr(), pc.bn.running_var.data_ptr(), BN_EPS, ctot,
                               scale.data_ptr(), shift.data_ptr())
            wf, bf = self._small((kp1, ctot)), self._small((kp1,))
            fold = self._op if self.training else self._setup_op
            fold("bpb_pixel_fold", pc.classifier.weight.data_ptr(), pc.classifier.bias.data_ptr(), scale.data_ptr(),
                 shift.data_ptr(), kp1, ctot, wf.data_ptr(), bf.data_ptr())
        self._alloc_head_outputs(h, w, ctot)
        mts = [self._small((n, J, b.h * b.w)) for b in br]
        if learn:
            self.fork(range(1, len(br)))
            for i, (b, off) in enumerate(zip(br, offs)):
                lo = self._small((n * b.h * b.w, kp1))
                with self.lane(i):
                    self._op("bpb_head_lowres_logits", b.t.data_ptr(), wf.data_ptr(), n * b.h * b.w, b.c, off, ctot, kp1,
                             lo.data_ptr())
                lows.append(lo)
            self.join()
        ext = self._ext_resized(h, w)
        self.mask_sum = self._small((n, J))
        self._op("bpb_head_pixel", self._ptr_array(lows) if learn else None, self._ptr_array(mts),
                 self._int_array([b.h for b in br]), self._int_array([b.w for b in br]), len(br), _lib.ptr(bf), n, h, w, kp1,
                 self.logits.data_ptr(), self.probs.data_ptr(), self.fg_mask.data_ptr(), self.vis_bits.data_ptr(),
                 self.vis_cont.data_ptr(), self.mask_sum.data_ptr(), _lib.ptr(ext), self.ext_mode)
        self._op("bpb_vis_unpack", self.vis_bits.data_ptr(), self.vis_cont.data_ptr(), n, kp1, self.vis_b.data_ptr(),
                 self.vis_f.data_ptr())
        self.fork(range(1, len(br)))
        for i, (b, off, mt) in enumerate(zip(br, offs, mts)):
            with self.lane(i):
                self._op("bpb_head_branch_pool", b.t.data_ptr(), mt.data_ptr(), self.mask_sum.data_ptr(), n, b.h * b.w,
                         b.c, off, ctot, kp1, h * w, self.pooled.data_ptr())
        self.join()
        self.hf = dict(br=br, offs=offs, ctot=ctot, scale=scale, shift=shift, wf=wf, bf=bf, lows=lows, mts=mts, h=h, w=w)
        return ctot

    def build_head_back(self, c):
        m, K, n = self.model, self.K, self.n
        J = K + 3
        D = m.dim_reduce_output
        ncls = m.num_classes
        ld = (ncls + 3) // 4 * 4
        # dim reduce: Linear(C->D) + BN1d + ReLU for {global, fg, bg, parts (K*N rows jointly)}
        lin = self._small((J * n, D))
        self.lin = lin
        self.emb = self._small((J * n, D))
        groups = [(m.global_after_pooling_dim_reduce, 0, n), (m.foreground_after_pooling_dim_reduce, n, n),
                  (m.background_after_pooling_dim_reduce, 2 * n, n), (m.parts_after_pooling_dim_reduce, 3 * n, K * n)]
        pooled2d = self.pooled.view(J * n, c)
        self.fork(range(1, len(groups)))
        for i, (dr, r0, rows) in enumerate(groups):
            fc, bn = dr.layers[0], dr.layers[1]
            with self.lane(i):
                self._op("bpb_linear_tf32", pooled2d[r0:].data_ptr(), c, fc.weight.data_ptr(), fc.bias.data_ptr(),
                         lin[r0:].data_ptr(), rows, D, c, D, 0, 0)
                self._bn1d(lin[r0:], bn, self.emb[r0:], rows, D, relu=True)
        self.join()
        # BN classifiers
        self.bn_emb = self._small((J * n, D))
        self.bn_conct = self._small((K * n, D))
        self.scores = self._small((J * n, ld))
        self.scores_conct = self._small((n, ld))
        heads = [(m.global_identity_classifier, 0), (m.foreground_identity_classifier, n),
                 (m.background_identity_classifier, 2 * n)]
        if m.shared_parts_id_classifier:
            heads_parts = [(m.parts_identity_classifier, 3 * n, K * n)]
        else:
            heads_parts = [(m.parts_identity_classifier[k], (3 + k) * n, n) for k in range(K)]
        # the identity heads are independent of each other: spread them over four stream lanes
        self.fork(range(1, 4))
        jobs = [(cl, r0, n) for cl, r0 in heads] + list(heads_parts)
        for i, (cl, r0, rows) in enumerate(jobs):
            with self.lane(i % 3):
                self._bn_classifier(cl, self.emb[r0:], self.bn_emb[r0:], self.scores[r0:], rows, D, ncls, ld)
        # concat-parts classifier: BN1d over [N, K*D] == per-part BN1d with the matching parameter slice
        cc = m.concat_parts_identity_classifier
        with self.lane(3):
            for k in range(K):
                sl = slice(k * D, (k + 1) * D)
                self._op("bpb_bn1d", self.emb[(3 + k) * n:].data_ptr(), cc.bn.weight[sl].data_ptr(),
                         cc.bn.bias[sl].data_ptr(), BN_EPS, 0.1, cc.bn.running_mean[sl].data_ptr(),
                         cc.bn.running_var[sl].data_ptr(), self.bn_conct[k * n:].data_ptr(), n, D, 0,
                         1 if self.training else 0)
            self._op("bpb_linear_tf32", self.bn_conct.data_ptr(), D, cc.classifier.weight.data_ptr(), None,
                     self.scores_conct.data_ptr(), n, ncls, K * D, ld, 1, K)
        self.join()
        self.ncls, self.ld, self.D = ncls, ld, D

    def _bn1d(self, x, bn, y, rows, c, relu):
        mom = bn.momentum if bn.momentum is not None else 0.1
        self._op("bpb_bn1d", x.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), BN_EPS, float(mom),
                 bn.running_mean.data_ptr(), bn.running_var.data_ptr(), y.data_ptr(), rows, c, 1 if relu else 0,
                 1 if self.training else 0)

    def _bn_classifier(self, cl, x, bn_out, scores, rows, D, ncls, ld):
        self._bn1d(x, cl.bn, bn_out, rows, D, relu=False)
        self._op("bpb_linear_tf32", bn_out.data_ptr(), D, cl.classifier.weight.data_ptr(), None, scores.data_ptr(),
                 rows, ncls, D, ld, 0, 0)

    # ------------------------------------------------------------------ build / run
    def _build(self):
        self.feats = None
        bb = self.model.backbone_appearance_feature_extractor
        if self.cfg.dim_reduce != "after_pooling" or self.cfg.pooling != "gwap" or self.cfg.normalization != "identity":
            raise NotImplementedError("only the shipped configuration (gwap / identity / after_pooling) is built")
        lowres = getattr(self.model, "lowres_head", True)
        want_map = getattr(self.model, "materialize_spatial_features", True)
        if isinstance(bb, P.HRNet32Params):
            br = self.build_hrnet(bb)
            if lowres:
                c = self.build_head_front_branches(br)
                if want_map:
                    self.feats = self.hrnet_concat(br)
            else:
                self.feats = self.hrnet_concat(br)
                self.build_head_front_map(self.feats)
                c = self.feats.c
        elif isinstance(bb, P.ResNet50Params):
            self.feats = self.build_resnet(bb)
            self.build_head_front_map(self.feats)
            c = self.feats.c
        else:
            raise KeyError(f"unknown backbone container {type(bb)}")
        self.build_head_back(c)

    def _run_list(self, ops):
        stream = torch.cuda.current_stream().cuda_stream
        err = self.lib.bpb_last_error
        for fn, args, name in ops:
            rc = fn(*args, stream)
            if rc != 0:
                raise RuntimeError(f"{name} failed (rc={rc}): {err().decode()}")

    def _launch_all(self):
        if self._need_setup:
            self._run_list(self.setup_ops)
            self._need_setup = False
        self._launch_lanes(self.ops)

    def _launch_lanes(self, ops):
        """launch a recorded op list: (fn, args, name, lane) entries on their stream lanes, (None, lanes, 'fork'|'join', 0)
        markers as stream dependencies"""
        main = torch.cuda.current_stream()
        nside = max((op[3] for op in ops), default=0)
        while len(self._side_streams) < nside:
            self._side_streams.append(torch.cuda.Stream(device=self.device))
        streams = [main] + self._side_streams
        ptrs = [st.cuda_stream for st in streams]
        active = []
        err = self.lib.bpb_last_error
        for fn, args, name, lane in ops:
            if fn is None:
                if name == "fork":
                    for l in args:
                        if l not in active:
                            streams[l].wait_stream(main)
                            active.append(l)
                else:
                    for l in active:
                        main.wait_stream(streams[l])
                    active = []
                continue
            if lane != 0 and lane not in active:             # op on a lane nobody forked: order it after lane 0
                streams[lane].wait_stream(main)
                active.append(lane)
            rc = fn(*args, ptrs[lane])
            if rc != 0:
                raise RuntimeError(f"{name} failed (rc={rc}): {err().decode()}")
        for l in active:
            main.wait_stream(streams[l])

    def run(self, images, external_parts_masks=None):
        if self.ext_mode:
            assert external_parts_masks is not None
            self.ext_in.copy_(external_parts_masks, non_blocking=True)
        ver = getattr(self.model, "_param_version", 0)
        if ver != getattr(self, "_seen_version", 0):          # an optimizer step changed the fp32 masters
            self._seen_version = ver
            if self.flat is None:
                self.refresh_weights()
        self.images.copy_(images, non_blocking=True)
        if self._need_setup:
            self._run_list(self.setup_ops)
            self._need_setup = False
        if self.use_graph:
            if self.graph is None:
                # warm-up outside capture (first-call attribute setup), then capture
                self._launch_all()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
        else:
            self._launch_all()

    # ------------------------------------------------------------------ reference-shaped outputs
    def outputs(self):
        """The 6-tuple of BPBreID.forward (bpbreid.py:222-259) as views of the plan's buffers."""
        n, K, D, ncls = self.n, self.K, self.D, self.ncls
        e, be, sc = self.emb, self.bn_emb, self.scores
        parts = e[3 * n:].view(K, n, D).permute(1, 0, 2)
        bn_parts = be[3 * n:].view(K, n, D).permute(1, 0, 2)
        binary = (self.training and self.cfg.training_binary_visibility_score) or \
                 (not self.training and self.cfg.testing_binary_visibility_score)
        vis = self.vis_b.view(torch.bool) if binary else self.vis_f
        embeddings = {
            "globl": e[0:n], "backg": e[2 * n:3 * n], "foreg": e[n:2 * n],
            "conct": parts.reshape(n, K * D), "parts": parts,
            "bn_globl": be[0:n], "bn_backg": be[2 * n:3 * n], "bn_foreg": be[n:2 * n],
            "bn_conct": self.bn_conct.view(K, n, D).permute(1, 0, 2).reshape(n, K * D), "bn_parts": bn_parts,
        }
        fg_v = vis[:, K + 1]
        visibility = {"globl": torch.ones_like(fg_v), "backg": vis[:, 0], "foreg": fg_v, "conct": fg_v,
                      "parts": vis[:, 1:K + 1]}
        scores = {"globl": sc[0:n, :ncls], "backg": sc[2 * n:3 * n, :ncls], "foreg": sc[n:2 * n, :ncls],
                  "conct": self.scores_conct[:, :ncls],
                  "parts": sc[3 * n:].view(K, n, self.ld).permute(1, 0, 2)[:, :, :ncls]}
        f = self.feats
        spatial = None if f is None else f.t.view(n, f.h, f.w, f.c).permute(0, 3, 1, 2)
        masks = {"globl": torch.ones_like(self.fg_mask), "backg": self.probs[:, 0], "foreg": self.fg_mask,
                 "conct": self.fg_mask, "parts": self.probs[:, 1:]}
        pix = self.logits if (self.ext_mode & 1) == 0 else None     # bpbreid.py:150: no pixel scores without the classifier
        return embeddings, visibility, scores, pix, spatial, masks
