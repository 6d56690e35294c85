"""Training step of the hot path on sm_100a: forward -> GiLt (+ body-part attention) loss -> ANALYTIC backward ->
(NCCL gradient all-reduce) -> Adam, as one replayable sequence of C-ABI launches over static buffers.

Reference call stack replaced (torchreid/engine/image/part_based_engine.py:77-105):
    model(imgs) ; combine_losses(...) ; optimizer.zero_grad() ; loss.backward() ; optimizer.step()
with torch.optim.Adam built by torchreid/optim/optimizer.py:96-103 (lr 3.5e-4, weight_decay 5e-4, betas 0.9/0.999;
default_config.py:125-131) and, on several GPUs, nn.DataParallel's gather / scatter of gradients (scripts/main.py:256-257)
replaced by one process per GPU with an NCCL all-reduce of the flat gradient buffer.

Design
  * FlatParams: every nn.Parameter of the model is re-pointed into ONE fp32 buffer (module attribute names and
    state_dict keys unchanged); a same-shaped gradient buffer and two Adam moment buffers sit beside it.  Conv weights
    that run on the tensor cores keep two bf16 operand copies (OHWI for the forward conv, [cin][flipped taps][cout]
    for the data gradient) refreshed by the same kernel that applies Adam (bpb_adam_pack).
  * TrainPlan(ForwardPlan): the forward planner in gradient mode (no buffer reuse, BatchNorm layers save mean/invstd).
    Every building block pushes a closure on a tape; building the backward walks the tape in reverse and records the
    backward launches:  ReLU / residual / fuse-sum gradient merge -> BatchNorm backward (reduce + apply) -> conv wgrad
    (tcgen05, MN-major) + conv dgrad (the forward tcgen05 kernel on dY with the flipped weight copy; stride-2 layers
    through a zero-inserted dY) ; head and loss backward in fp32.
  * TrainStep: owns plan + loss + optimizer state; `step(images, pids, masks)` runs everything (CUDA graph when
    single-GPU), returns the loss tensor.  `loss.backward()`-style use from foreign code goes through
    `bpbreid_b200.autograd_bridge`.
Nothing here computes numbers in Python: every value is produced by libbpbreid_sm100.so.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib
from . import params as P
from .constants import GLOBAL, FOREGROUND, BACKGROUND, CONCAT_PARTS, PARTS
from .plan import BN_EPS, Act, ForwardPlan, Lazy

ALIGN = 64          # floats: every parameter segment starts on a 256-byte boundary (TMA / vector-reduction alignment)
SCORE_LD = 768      # padded class dimension of the score gradients (multiple of 32 >= num_classes)


class FlatParams:
    """One flat fp32 buffer for all parameters (+ gradient and Adam moments), segment table for bpb_adam_pack."""

    def __init__(self, model: nn.Module, device):
        self.model, self.device = model, device
        self.entries = []                      # (name, param, off, numel)
        off = 0
        for name, p in model.named_parameters():
            self.entries.append([name, p, off, p.numel()])
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.p = torch.zeros(off, device=device, dtype=torch.float32)
        self.g = torch.zeros(off, device=device, dtype=torch.float32)
        self.m = torch.zeros(off, device=device, dtype=torch.float32)
        self.v = torch.zeros(off, device=device, dtype=torch.float32)
        self.off = {}
        with torch.no_grad():
            for name, p, o, n in self.entries:
                self.p[o:o + n].copy_(p.detach().reshape(-1).to(device))
                p.data = self.p[o:o + n].view(p.shape)
                self.off[id(p)] = o
        self.conv_ops = {}                     # id(conv module) -> (wf bf16 [cout, k*k*cin], wd bf16 [cin, k*k*cout])
        self.touched = set()                   # id(param) of parameters that receive a gradient
        self.seg_dev = None
        self.step_count = 0
        self.bounds = {}                       # bucket name -> (lo, hi) flat range

    # -- views ------------------------------------------------------------------------------------------------
    def grad_ptr(self, p, elem_off=0):
        self.touched.add(id(p))
        return self.g.data_ptr() + 4 * (self.off[id(p)] + elem_off)

    def grad_view(self, p):
        o = self.off[id(p)]
        return self.g[o:o + p.numel()].view(p.shape)

    def conv_grad_view(self, conv):
        """Gradient of a tensor-core conv weight as the reference-shaped [cout, cin, k, k] (strided) view of the OHWI
        gradient the wgrad kernel accumulates."""
        w = conv.weight
        o = self.off[id(w)]
        co, ci, k, _ = w.shape
        return self.g[o:o + w.numel()].view(co, k, k, ci).permute(0, 3, 1, 2)

    def forward_operand(self, conv):
        return self._operands(conv)[0]

    def dgrad_operand(self, conv):
        return self._operands(conv)[1]

    def _operands(self, conv):
        key = id(conv)
        if key not in self.conv_ops:
            co, ci, k, _ = conv.weight.shape
            wf = torch.zeros(co, k * k * ci, device=self.device, dtype=torch.bfloat16)
            wd = torch.zeros(ci, k * k * co, device=self.device, dtype=torch.bfloat16)
            self.conv_ops[key] = (wf, wd, conv)
            self.seg_dev = None
        return self.conv_ops[key]

    # -- segment table ----------------------------------------------------------------------------------------
    def build_segments(self):
        seg_bytes = _lib.query("bpb_param_segment_bytes")
        dt = np.dtype([("off", "<i8"), ("numel", "<i8"), ("kind", "<i4"), ("trainable", "<i4"), ("cout", "<i4"),
                       ("cin", "<i4"), ("k", "<i4"), ("pad", "<i4"), ("wf", "<u8"), ("wd", "<u8")])
        assert dt.itemsize == seg_bytes, (dt.itemsize, seg_bytes)
        conv_of = {id(c.weight): (wf, wd, c) for wf, wd, c in self.conv_ops.values()}
        # padding between segments belongs to the preceding segment (numel only bounds the binary search)
        arr = np.zeros(len(self.entries), dtype=dt)
        for i, (name, p, o, n) in enumerate(self.entries):
            arr[i]["off"], arr[i]["numel"] = o, n
            arr[i]["trainable"] = 1 if (p.requires_grad and id(p) in self.touched) else 0
            if id(p) in conv_of:
                wf, wd, c = conv_of[id(p)]
                co, ci, k, _ = p.shape
                arr[i]["kind"], arr[i]["cout"], arr[i]["cin"], arr[i]["k"] = 1, co, ci, k
                arr[i]["wf"], arr[i]["wd"] = wf.data_ptr(), wd.data_ptr()
        self.seg_host = arr
        self.seg_dev = torch.from_numpy(arr.view(np.uint8).copy()).to(self.device)
        self.nseg = len(self.entries)

    def pack_args(self, lr=0.0, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, step=1, grad_scale=1.0, do_update=0):
        if self.seg_dev is None:
            self.build_segments()
        return ("bpb_adam_pack", self.seg_dev.data_ptr(), self.nseg, self.total, self.p.data_ptr(), self.g.data_ptr(),
                self.m.data_ptr(), self.v.data_ptr(), float(lr), float(beta1), float(beta2), float(eps), float(wd),
                int(step), float(grad_scale), int(do_update))

    def pack_now(self):
        """Refresh the bf16 operand copies from the fp32 masters (after load_state_dict / at build)."""
        a = self.pack_args()
        _lib.call(a[0], *a[1:], _lib.stream_ptr())


class TrainPlan(ForwardPlan):
    """ForwardPlan in gradient mode + the recorded backward pass."""

    def __init__(self, model, n, h, w, device, flat: FlatParams, loss_cfg, debug=False):
        self.loss_cfg = loss_cfg
        self.bops = []                 # backward launches: (fn, args, name)
        self._gl = {}                  # data_ptr of an activation -> list of gradient Acts
        self._gfree = {}
        self._gref = {}
        self._gall = []
        self.bwd_ws = {}
        self.btrace = []               # debug: one record per conv+BN backward (teacher-forced kernel checks in tests)
        self.head_dx = []              # debug: gradients w.r.t. the branch maps as the head backward leaves them
        self.marks = []                # (index into bops, stage name): gradient-bucket boundaries for the all-reduce
        self._blane = 0                # stream lane the backward launches are being recorded on
        super().__init__(model, n, h, w, True, device, "tc", False, debug, grad=True, flat=flat)

    # ------------------------------------------------------------------ gradient bookkeeping
    # Gradient buffers are plain device tensors whose raw pointers are baked into the recorded launches, so the plan
    # must own them for its whole life (self._gall); a buffer goes back to the free list -- and may be handed out
    # again -- only once every launch that reads it has been RECORDED (launches replay in recording order on one
    # stream, so a later writer can never overtake an earlier reader).
    def g_add(self, act: Act, g: Act):
        self._gl.setdefault(act.t.data_ptr(), []).append(g)
        self._gref[id(g.buf)] += 1

    def g_take(self, act: Act):
        return self._gl.pop(act.t.data_ptr(), [])

    def galloc(self, n, h, w, c) -> Act:
        nbytes = n * h * w * c * 2
        lst = self._gfree.get((self._blane, nbytes))
        if lst:
            buf = lst.pop()
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._gall.append(buf)
        self._gref[id(buf)] = 0
        return Act(buf.view(torch.bfloat16).view(n, h, w, c), n, h, w, c, buf)

    def g_release(self, g: Act):
        """one consumer of `g` has been recorded"""
        self._gref[id(g.buf)] -= 1
        self.g_done(g)

    def g_done(self, g: Act):
        """return `g` to the pool if nobody else is going to read it (never in debug mode: tests inspect the buffers)"""
        if self._gref[id(g.buf)] <= 0 and not self.debug:
            self._gref[id(g.buf)] = 1 << 30                    # guards against a double free
            # the buffer joins the pool of the lane whose launch read it LAST: the next writer is then ordered behind
            # that reader on the same stream (the producer's lane may be running concurrently)
            self._gfree.setdefault((self._blane, g.buf.numel()), []).append(g.buf)

    def _bop(self, name, *args):
        self.bops.append((getattr(self.lib, name), args, name, self._blane))

    def _bn_ws(self, c):
        key = (self._blane, c)                                  # concurrent lanes must not share the partial-sum rows
        if key not in self.bwd_ws:
            self.bwd_ws[key] = torch.zeros(_lib.query("bpb_bn_bwd_workspace_bytes", c), dtype=torch.uint8,
                                           device=self.device)
        return self.bwd_ws[key]

    def grad_sum(self, grads, y: Act | None):
        """g = (sum grads) * [y > 0]"""
        like = grads[0]
        partials = []
        while len(grads) > 4:                                   # (never happens on the shipped topologies)
            part = self.galloc(like.n, like.h, like.w, like.c)
            self._bop("bpb_relu_grad_sum", self._ptr_array([g.t for g in grads[:4]]), 4, None, part.t.data_ptr(),
                      like.m, like.c)
            grads = [part] + list(grads[4:])
            partials.append(part)
        out = self.galloc(like.n, like.h, like.w, like.c)
        self._bop("bpb_relu_grad_sum", self._ptr_array([g.t for g in grads]), len(grads),
                  None if y is None else y.t.data_ptr(), out.t.data_ptr(), like.m, like.c)
        for part in partials:
            self.g_done(part)
        return out

    # ------------------------------------------------------------------ backward of the conv + BN pair
    def lazy_backward(self, lz: Lazy, dy: Act, relu_affine: int, up: int):
        raw, conv, bn = lz.raw, lz.conv, lz.bn
        assert conv is not None and lz.save is not None, "Lazy without backward information"
        fl = self.flat
        coef = self._small((2, raw.c))
        self._bop("bpb_bn_bwd_reduce", dy.t.data_ptr(), raw.t.data_ptr(), lz.save.data_ptr(), lz.scale.data_ptr(),
                  lz.shift.data_ptr(), raw.n, raw.h, raw.w, raw.c, up, relu_affine, 0, self._bn_ws(raw.c).data_ptr(),
                  coef.data_ptr(), fl.grad_ptr(bn.weight), fl.grad_ptr(bn.bias))
        stride = conv.stride[0]
        zero_ins = 1 if (stride == 2 and lz.kind == "tc") else 0
        draw = self.galloc(raw.n, raw.h * (2 if zero_ins else 1), raw.w * (2 if zero_ins else 1), raw.c)
        self._bop("bpb_bn_bwd_apply", dy.t.data_ptr(), raw.t.data_ptr(), lz.save.data_ptr(), lz.scale.data_ptr(),
                  lz.shift.data_ptr(), coef.data_ptr(), raw.n, raw.h, raw.w, raw.c, up, relu_affine, zero_ins,
                  draw.t.data_ptr())
        k = conv.kernel_size[0]
        if self.debug:
            self.btrace.append(dict(lz=lz, dy=dy, relu_affine=relu_affine, up=up, draw=draw, zero_ins=zero_ins, dx=None,
                                    name=self._names.get(id(conv), "?")))
        if lz.kind == "stem":
            self._bop("bpb_stem_wgrad", self.images.data_ptr(), draw.t.data_ptr(), fl.grad_ptr(conv.weight), self.n,
                      self.h, self.w, conv.out_channels, k, stride)
            self.g_done(draw)
            return
        x = lz.x_in
        cin, cout = conv.in_channels, conv.out_channels
        self._bop("bpb_conv_wgrad_bf16", x.t.data_ptr(), draw.t.data_ptr(), fl.grad_ptr(conv.weight), x.n, x.h, x.w,
                  cin, cout, k, stride, zero_ins)
        wd = fl.dgrad_operand(conv)
        dx = self.galloc(x.n, x.h, x.w, x.c)
        if zero_ins:
            assert draw.h == x.h and draw.w == x.w, "stride-2 dgrad needs even input sizes"
        # data gradient = the forward tensor-core conv on dY with the flipped / transposed weight copy, stride 1
        self._bop("bpb_conv_bf16", draw.t.data_ptr(), wd.data_ptr(), dx.t.data_ptr(), None, draw.n, draw.h, draw.w,
                  cout, cin, k, 1, 0)
        if self.debug:
            self.btrace[-1]["dx"] = dx
        self.g_add(x, dx)
        self.g_done(draw)

    # ------------------------------------------------------------------ taped building blocks
    def apply(self, lz: Lazy, relu, res=None) -> Act:
        y = super().apply(lz, relu, res)

        def bwd():
            grads = self.g_take(y)
            if not grads:
                return
            if relu and res is None and len(grads) == 1:
                dy, relu_affine = grads[0], 1                   # ReLU gate recomputed from raw*scale+shift
            else:
                dy = grads[0] if (len(grads) == 1 and not relu) else self.grad_sum(grads, y if relu else None)
                relu_affine = 0
            self.lazy_backward(lz, dy, relu_affine, 0)
            if res is not None:
                if isinstance(res, Lazy):
                    self.lazy_backward(res, dy, 0, 0)
                else:
                    self.g_add(res, dy)
            fresh = all(dy is not g for g in grads)
            for g in grads:
                self.g_release(g)
            if fresh:
                self.g_done(dy)
        self.tape.append((bwd, self._lane))
        return y

    def fuse(self, terms, out_like: Act, relu=True) -> Act:
        y = super().fuse(terms, out_like, relu)

        def bwd():
            grads = self.g_take(y)
            if not grads:
                return
            g = self.grad_sum(grads, y if relu else None)
            for t, up in terms:
                if isinstance(t, Lazy):
                    self.lazy_backward(t, g, 0, up)
                else:
                    assert up == 0
                    self.g_add(t, g)
            for gg in grads:
                self.g_release(gg)
            self.g_done(g)
        self.tape.append((bwd, self._lane))
        return y

    def mark(self, name):
        self.tape.append((lambda: self.marks.append((len(self.bops), name)), 0))

    def ibn_relu(self, x, conv, ibn):
        raise NotImplementedError("training backward of the IBN backbone is not built (SURVEY.md section 8 config 4)")

    # ------------------------------------------------------------------ head (branch maps) backward
    def build_head_front_branches(self, br):
        ctot = super().build_head_front_branches(br)
        hf = self.hf
        n, kp1, J = self.n, self.K + 1, self.K + 3
        h, w = hf["h"], hf["w"]
        pc = self.model.pixel_classifier
        fl = self.flat
        self.g_pooled = self._small((J, n, ctot), zero=True)        # dL/dpooled, written by the dim-reduce backward
        self.target_masks = None
        self.pix_coef = 0.0

        def bwd():
            dms = self._small((n, J))
            self._bop("bpb_head_pool_bwd_prep", self.g_pooled.data_ptr(), self.pooled.data_ptr(),
                      self.mask_sum.data_ptr(), n, ctot, kp1, h * w, dms.data_ptr())
            dmts, dlows = [], []
            for b, off in zip(br, hf["offs"]):
                dmt = self._small((n, b.h * b.w, J))
                self._bop("bpb_head_dmask", b.t.data_ptr(), self.g_pooled.data_ptr(), n, b.h * b.w, b.c, off, ctot, kp1,
                          dmt.data_ptr())
                dmts.append(dmt)
                dlows.append(self._small((n * b.h * b.w, kp1)))
            dwf, dbf = self._small((kp1, ctot)), self._small((kp1,))
            self._bop("bpb_memset_zero", dwf.data_ptr(), dwf.numel() * 4)
            self._bop("bpb_memset_zero", dbf.data_ptr(), dbf.numel() * 4)
            tm = self.loss_cfg.get("bpa_masks")
            self._bop("bpb_head_pixel_bwd", self._ptr_array(dmts), self._ptr_array(dlows),
                      self._int_array([b.h for b in br]), self._int_array([b.w for b in br]), len(br),
                      self.probs.data_ptr(), dms.data_ptr(), n, h, w, kp1, None if tm is None else tm.data_ptr(),
                      0 if tm is None else tm.shape[2], 0 if tm is None else tm.shape[3],
                      float(self.loss_cfg.get("bpa_label_smoothing", 0.1)),
                      float(self.loss_cfg.get("bpa_weight", 0.0)) / float(n * h * w), None, dbf.data_ptr())
            for b, off, dl in zip(br, hf["offs"], dlows):
                self._bop("bpb_head_dwf", b.t.data_ptr(), dl.data_ptr(), n, b.h * b.w, b.c, off, ctot, kp1,
                          dwf.data_ptr())
            coefab = self._small((2, ctot))
            self._bop("bpb_pixel_fold_bwd", dwf.data_ptr(), dbf.data_ptr(), pc.classifier.weight.data_ptr(),
                      pc.bn.weight.data_ptr(), hf["scale"].data_ptr(), hf["shift"].data_ptr(),
                      self.pix_save.data_ptr(), kp1, ctot, n * h * w, fl.grad_ptr(pc.classifier.weight),
                      fl.grad_ptr(pc.classifier.bias), fl.grad_ptr(pc.bn.weight), fl.grad_ptr(pc.bn.bias),
                      coefab.data_ptr())
            for b, off, mt, dl in zip(br, hf["offs"], hf["mts"], dlows):
                dx = self.galloc(b.n, b.h, b.w, b.c)
                self._bop("bpb_head_branch_bwd", b.t.data_ptr(), mt.data_ptr(), self.g_pooled.data_ptr(), dl.data_ptr(),
                          hf["wf"].data_ptr(), coefab.data_ptr(), n, b.h, b.w, b.c, off, ctot, h, w, kp1,
                          dx.t.data_ptr())
                self.g_add(b, dx)
                if self.debug:
                    self.head_dx.append(dx)
            if self.debug:
                self.head_dbg = dict(dms=dms, dmts=dmts, dlows=dlows, dwf=dwf, dbf=dbf, coefab=coefab)
        self.tape.append((bwd, 0))
        return ctot

    def build_head_front_map(self, feats):
        raise NotImplementedError("training backward of the map-based head (ResNet backbones) is not built")

    # ------------------------------------------------------------------ head back (Linear / BN1d / classifiers)
    def _linear_bwd(self, dy, ld_dy, rows, nout, x, ld_x, nin, weight, dx, ld_dx, dw_ptr):
        """y[rows, nout] = x[rows, nin] W[nout, nin]^T:  dx = dy W ;  dW = dy^T x   (tf32 tcgen05 GEMMs; operands are
        staged K-major by bpb_transpose_pad).  dy may have a padded leading dimension (zero columns beyond nout)."""
        kpad = ld_dy                                        # contraction length of the dx GEMM (zero padded)
        rows_p = (rows + 3) // 4 * 4
        if dx is not None:
            wt = self._small((nin, kpad))                   # W^T, zero padded to kpad columns
            self._bop("bpb_transpose_pad", weight.data_ptr(), nin, nout, nin, wt.data_ptr(), kpad, nin, kpad)
            self._bop("bpb_linear_tf32", dy.data_ptr(), ld_dy, wt.data_ptr(), None, dx.data_ptr(), rows, nin, kpad,
                      ld_dx, 0, 0)
        dyt = self._small((nout, rows_p))
        xt = self._small((nin, rows_p))
        self._bop("bpb_transpose_pad", dy.data_ptr(), ld_dy, rows, nout, dyt.data_ptr(), rows_p, nout, rows_p)
        self._bop("bpb_transpose_pad", x.data_ptr(), ld_x, rows, nin, xt.data_ptr(), rows_p, nin, rows_p)
        self._bop("bpb_linear_tf32", dyt.data_ptr(), rows_p, xt.data_ptr(), None, dw_ptr, nout, nin, rows_p, nin, 0, 0)

    def build_head_back(self, c):
        super().build_head_back(c)
        m, K, n = self.model, self.K, self.n
        J, D, ncls = K + 3, self.D, self.ncls
        fl = self.flat
        lw = self.loss_cfg["weights"]
        self.d_emb = self._small((J * n, D), zero=True)
        self.d_scores = self._small((J * n, SCORE_LD), zero=True)
        self.d_scores_conct = self._small((n, SCORE_LD), zero=True)
        assert ncls <= SCORE_LD

        def bwd():
            # ---- BN classifiers with a cross-entropy term (GiLt_loss.py:53-66): globl, foreg rows of `scores`
            heads = [(GLOBAL, m.global_identity_classifier, 0, n), (FOREGROUND, m.foreground_identity_classifier, n, n)]
            if lw[PARTS]["id"] > 0:
                if m.shared_parts_id_classifier:
                    heads.append((PARTS, m.parts_identity_classifier, 3 * n, K * n))
                else:
                    heads += [(PARTS, m.parts_identity_classifier[k], (3 + k) * n, n) for k in range(K)]
            for key, cl, r0, rows in heads:
                if lw[key]["id"] <= 0:
                    continue
                dbn = self._small((rows, D))
                self._linear_bwd(self.d_scores[r0:], SCORE_LD, rows, ncls, self.bn_emb[r0:], D, D, cl.classifier.weight,
                                 dbn, D, fl.grad_ptr(cl.classifier.weight))
                self._bop("bpb_bn1d_bwd", dbn.data_ptr(), D, self.emb[r0:].data_ptr(), D, None, 0,
                          cl.bn.weight.data_ptr(), BN_EPS, rows, D, self.d_emb[r0:].data_ptr(), D, 1,
                          fl.grad_ptr(cl.bn.weight), None)
            # ---- concat-parts classifier: BN1d over [N, K*D] + Linear(K*D -> ID)
            if lw[CONCAT_PARTS]["id"] > 0:
                cc = m.concat_parts_identity_classifier
                KD = K * D
                rows_p = (n + 3) // 4 * 4
                dbnc = self._small((n, KD))
                wt = self._small((KD, SCORE_LD))
                self._bop("bpb_transpose_pad", cc.classifier.weight.data_ptr(), KD, ncls, KD, wt.data_ptr(), SCORE_LD, KD,
                          SCORE_LD)
                self._bop("bpb_linear_tf32", self.d_scores_conct.data_ptr(), SCORE_LD, wt.data_ptr(), None,
                          dbnc.data_ptr(), n, KD, SCORE_LD, KD, 0, 0)
                dyt = self._small((ncls, rows_p))
                xt = self._small((KD, rows_p))
                self._bop("bpb_transpose_pad", self.d_scores_conct.data_ptr(), SCORE_LD, n, ncls, dyt.data_ptr(), rows_p,
                          ncls, rows_p)
                for k in range(K):                           # bn_conct is part-major [K, n, D]
                    self._bop("bpb_transpose_pad", self.bn_conct[k * n:].data_ptr(), D, n, D, xt[k * D:].data_ptr(),
                              rows_p, D, rows_p)
                self._bop("bpb_linear_tf32", dyt.data_ptr(), rows_p, xt.data_ptr(), None,
                          fl.grad_ptr(cc.classifier.weight), ncls, KD, rows_p, KD, 0, 0)
                for k in range(K):
                    self._bop("bpb_bn1d_bwd", dbnc[:, k * D:].data_ptr(), KD, self.emb[(3 + k) * n:].data_ptr(), D, None,
                              0, cc.bn.weight[k * D:].data_ptr(), BN_EPS, n, D, self.d_emb[(3 + k) * n:].data_ptr(), D, 1,
                              fl.grad_ptr(cc.bn.weight, k * D), None)
            # ---- dim-reduce layers (Linear + BN1d + ReLU) of the groups that received a gradient
            groups = [(GLOBAL, m.global_after_pooling_dim_reduce, 0, n),
                      (FOREGROUND, m.foreground_after_pooling_dim_reduce, n, n),
                      (BACKGROUND, m.background_after_pooling_dim_reduce, 2 * n, n),
                      (PARTS, m.parts_after_pooling_dim_reduce, 3 * n, K * n)]
            gp2d = self.g_pooled.view(J * n, c)
            pooled2d = self.pooled.view(J * n, c)
            for key, dr, r0, rows in groups:
                live = (key != BACKGROUND and (lw[key]["id"] > 0 or lw[key]["tr"] > 0)) or \
                       (key == PARTS and lw[CONCAT_PARTS]["id"] > 0)
                if not live:
                    continue
                fc, bn = dr.layers[0], dr.layers[1]
                dlin = self._small((rows, D))
                self._bop("bpb_bn1d_bwd", self.d_emb[r0:].data_ptr(), D, self.lin[r0:].data_ptr(), D,
                          self.emb[r0:].data_ptr(), D, bn.weight.data_ptr(), BN_EPS, rows, D, dlin.data_ptr(), D, 0,
                          fl.grad_ptr(bn.weight), fl.grad_ptr(bn.bias))
                self._bop("bpb_colsum_f32", dlin.data_ptr(), D, rows, D, fl.grad_ptr(fc.bias))
                self._linear_bwd(dlin, D, rows, D, pooled2d[r0:], c, c, fc.weight, gp2d[r0:], c, fl.grad_ptr(fc.weight))
        self.tape.append((bwd, 0))

    # ------------------------------------------------------------------ build the backward op list
    def build_backward(self):
        """Walk the tape backwards.  Entries carry the stream lane they were recorded on in the forward pass; the
        forward's fork / join markers turn into join / fork: the branches of an HRNet module (and its fuse outputs) are
        independent in the backward pass exactly where they were in the forward pass."""
        assert not self._gl
        lanes_on = self.multi_stream and not self.debug
        for entry in reversed(self.tape):
            if isinstance(entry[0], str):
                kind, lanes = entry
                if lanes_on:
                    # reversed region: opens at the forward's join (with the lanes that region used), closes at its fork
                    self.bops.append((None, tuple(lanes), "fork", 0) if kind == "join" else (None, (), "join", 0))
                continue
            fn, lane = entry
            self._blane = lane if lanes_on else 0
            fn()
        self._blane = 0
        self._gl.clear()


class TrainStep:
    """One fused training step (forward, loss, backward, all-reduce, Adam) on static buffers."""

    def __init__(self, model, criterion=None, optimizer=None, bpa_weight=0.0, lr=3.5e-4, weight_decay=5e-4,
                 betas=(0.9, 0.999), eps=1e-8, use_graph=True, group=None):
        from . import losses
        self.model = model
        self.crit = criterion if criterion is not None else losses.GiLtLoss(None, use_gpu=True)
        if optimizer is not None:                              # take the hyper-parameters of a torch.optim.Adam
            g0 = optimizer.param_groups[0]
            lr, weight_decay, betas, eps = g0["lr"], g0.get("weight_decay", 0.0), g0.get("betas", betas), g0.get("eps", eps)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.bpa_weight = float(bpa_weight)
        self.use_graph = use_graph
        self.group = group
        self.debug = False               # keep per-layer backward records (plan.btrace) for the kernel parity tests
        self.flat = None
        self.plans = {}
        self.graphs = {}

    # ------------------------------------------------------------------ plan construction
    def _plan(self, images, masks):
        n, _, h, w = images.shape
        key = (n, h, w, None if masks is None else tuple(masks.shape))
        if key in self.plans:
            return self.plans[key]
        dev = images.device
        if self.flat is None:
            self.flat = FlatParams(self.model, dev)
        self.model._plans = {}                                   # parameters moved: cached forward plans are stale
        lw = {k: dict(v) for k, v in self.crit.losses_weights.items() if isinstance(v, dict) and "id" in v}
        for k in (GLOBAL, FOREGROUND, CONCAT_PARTS, PARTS):
            lw.setdefault(k, {"id": 0.0, "tr": 0.0})
        lw.setdefault(BACKGROUND, {"id": 0.0, "tr": 0.0})
        cfg = dict(weights=lw, bpa_weight=self.bpa_weight if masks is not None else 0.0, bpa_label_smoothing=0.1,
                   bpa_masks=None if masks is None else torch.zeros_like(masks, dtype=torch.float32, device=dev))
        self.model.train()
        plan = TrainPlan(self.model, n, h, w, dev, self.flat, cfg, debug=self.debug)
        plan.masks_buf = cfg["bpa_masks"]
        self._build_loss(plan, lw)
        plan.build_backward()
        self.flat.build_segments()
        self.flat.pack_now()
        self.plans[key] = plan
        return plan

    def _build_loss(self, plan, lw):
        """GiLt forward + gradient seeds as recorded launches (GiLt_loss.py:45-119)."""
        n, K, D, ncls, ld = plan.n, plan.K, plan.D, plan.ncls, plan.ld
        dev = plan.device
        plan.pids = torch.zeros(n, dtype=torch.int64, device=dev)
        plan.loss_ops = []
        terms, weights = [], []
        world = self._world()

        def lop(name, *args):
            plan.loss_ops.append((getattr(plan.lib, name), args, name))
        vis_on = bool(getattr(self.crit, "use_visibility_scores", False))
        if vis_on:
            raise NotImplementedError("mask_filtering_training=True (visibility-weighted losses) in the fused training "
                                      "step is not built; use the forward-only losses")
        margin = float(self.crit.part_triplet_loss.margin)
        eps_ls = float(self.crit.identity_loss.eps)
        ce_heads = [(GLOBAL, plan.scores[0:], plan.d_scores[0:], n), (FOREGROUND, plan.scores[n:], plan.d_scores[n:], n),
                    (CONCAT_PARTS, plan.scores_conct, plan.d_scores_conct, n)]
        if lw[PARTS]["id"] > 0:
            # [N, K, ID] flattened == K*n rows of the part-major score buffer with the labels repeated per part
            plan.pids_rep = torch.zeros(K * n, dtype=torch.int64, device=dev)
            ce_heads.append((PARTS, plan.scores[3 * n:], plan.d_scores[3 * n:], K * n))
        for key, sc, dsc, rows in ce_heads:
            wgt = float(lw[key]["id"])
            if wgt <= 0:
                continue
            scratch = torch.empty(2 * rows, device=dev)
            out2 = torch.empty(2, device=dev)
            tgt = plan.pids if rows == n else plan.pids_rep
            lop("bpb_ce_ls_fwd", sc.data_ptr(), ld, tgt.data_ptr(), None, rows, ncls, eps_ls, 0, scratch.data_ptr(),
                out2.data_ptr())
            # identity terms are per-sample means: local mean / world == the global-batch mean after the gradient sum
            lop("bpb_ce_ls_bwd", sc.data_ptr(), ld, tgt.data_ptr(), None, rows, ncls, eps_ls, wgt / world,
                dsc.data_ptr(), SCORE_LD)
            plan.persistent += [scratch, out2]
            terms.append(out2)
            weights.append(wgt)
        plan.tri = []
        tri_keys = [(PARTS, plan.emb[3 * n:], plan.d_emb[3 * n:], K), (FOREGROUND, plan.emb[n:], plan.d_emb[n:], 1),
                    (GLOBAL, plan.emb[0:], plan.d_emb[0:], 1)]
        plan.sharded = []                                        # (ShardedTriplet, local emb view, local grad view, weight)
        for key, e, de, kk in tri_keys:
            wgt = float(lw[key]["tr"])
            if wgt <= 0:
                continue
            if world > 1:
                # global-batch mining with per-rank anchor rows (bpbreid_b200.distributed.ShardedTriplet): the NCCL calls
                # sit between recorded launch segments, see _launch
                from .distributed import ShardedTriplet
                st = ShardedTriplet(n, kk, D, margin, dev, self.group)
                plan.sharded.append((st, e[:kk * n].view(kk, n, D), de[:kk * n].view(kk, n, D), wgt))
                terms.append(st.out4)
                weights.append(wgt)
                continue
            bufs = dict(dist=torch.empty(n, n, device=dev), hp=torch.empty(n, device=dev), hn=torch.empty(n, device=dev),
                        hp_idx=torch.empty(n, device=dev, dtype=torch.int32),
                        hn_idx=torch.empty(n, device=dev, dtype=torch.int32), out4=torch.empty(4, device=dev))
            lop("bpb_triplet_fwd", e.data_ptr(), plan.pids.data_ptr(), None, 0, n, kk, D, margin, bufs["dist"].data_ptr(),
                bufs["hp"].data_ptr(), bufs["hn"].data_ptr(), bufs["hp_idx"].data_ptr(), bufs["hn_idx"].data_ptr(),
                bufs["out4"].data_ptr())
            lop("bpb_triplet_bwd", e.data_ptr(), None, 0, bufs["hp"].data_ptr(), bufs["hn"].data_ptr(),
                bufs["hp_idx"].data_ptr(), bufs["hn_idx"].data_ptr(), bufs["out4"].data_ptr(), n, kk, D, margin,
                wgt / world, de.data_ptr())
            plan.tri.append((key, bufs))
            terms.append(bufs["out4"])
            weights.append(wgt)
        if lw[CONCAT_PARTS]["tr"] > 0:
            raise NotImplementedError("triplet term on the concatenated parts embedding is not built")
        if plan.masks_buf is not None and self.bpa_weight > 0:
            partial = torch.empty(512, device=dev)
            out2 = torch.empty(2, device=dev)
            tm = plan.masks_buf
            lop("bpb_pixel_ce_fwd", plan.logits.data_ptr(), tm.data_ptr(), n, K + 1, plan.logits.shape[2],
                plan.logits.shape[3], tm.shape[2], tm.shape[3], 0.1, partial.data_ptr(), out2.data_ptr())
            plan.persistent += [partial, out2]
            terms.append(out2)
            weights.append(self.bpa_weight)
        plan.loss_out = torch.zeros(1, device=dev)
        plan.loss_terms = terms
        ptrs = (ctypes.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
        ws = (ctypes.c_float * len(terms))(*weights)
        plan.keepalive += [ptrs, ws]
        lop("bpb_weighted_sum", ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(ws, ctypes.c_void_p), len(terms),
            plan.loss_out.data_ptr())

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    # ------------------------------------------------------------------ execution
    def _launch(self, plan, update=True):
        fl = self.flat
        stream = torch.cuda.current_stream().cuda_stream
        err = plan.lib.bpb_last_error

        def run(ops):
            for fn, args, name in ops:
                rc = fn(*args, stream)
                if rc != 0:
                    raise RuntimeError(f"{name} failed (rc={rc}): {err().decode()}")
        if self._world() == 1:
            plan._launch_all()                                  # forward (multi-lane)
            # gradient buffers that are accumulated into
            plan.lib.bpb_memset_zero(fl.g.data_ptr(), fl.total * 4, stream)
            plan.lib.bpb_memset_zero(plan.d_emb.data_ptr(), plan.d_emb.numel() * 4, stream)
            run(plan.loss_ops)
            plan._launch_lanes(plan.bops)
            return
        for kind, fn in self._phases(plan):
            fn()

    def _phases(self, plan):
        """Data-parallel step as alternating kernel phases ('g': recorded launches only -- each one is captured in its own
        CUDA graph by forward_backward) and NCCL phases ('n').  The backward is cut at the stage boundaries; the gradient
        bucket that is complete after a segment is all-reduced on the NCCL stream while the next segment runs."""
        import torch.distributed as dist
        fl = self.flat

        def run(ops):
            if ops and len(ops[0]) == 4:
                plan._launch_lanes(ops)
                return
            stream = torch.cuda.current_stream().cuda_stream
            for fn, args, name in ops:
                rc = fn(*args, stream)
                if rc != 0:
                    raise RuntimeError(f"{name} failed (rc={rc}): {plan.lib.bpb_last_error().decode()}")

        def fwd():
            stream = torch.cuda.current_stream().cuda_stream
            plan._launch_all()
            plan.lib.bpb_memset_zero(fl.g.data_ptr(), fl.total * 4, stream)
            plan.lib.bpb_memset_zero(plan.d_emb.data_ptr(), plan.d_emb.numel() * 4, stream)
        works = []
        ph = [("g", fwd),
              ("n", lambda: [st.gather(e, plan.pids) for st, e, de, w in plan.sharded]),
              ("g", lambda: [st.mine() for st, e, de, w in plan.sharded]),
              ("n", lambda: [st.reduce_stats() for st, e, de, w in plan.sharded]),
              ("g", lambda: ([st.finish() for st, e, de, w in plan.sharded], run(plan.loss_ops),
                             [st.bwd_kernel(w) for st, e, de, w in plan.sharded])),
              ("n", lambda: [st.bwd_reduce() for st, e, de, w in plan.sharded])]
        lo_i = 0
        first = True
        for hi_i, (g_lo, g_hi) in self._buckets(plan):
            seg = plan.bops[lo_i:hi_i]
            if first:
                ph.append(("g", lambda seg=seg: ([st.bwd_copy(de) for st, e, de, w in plan.sharded], run(seg))))
                first = False
            else:
                ph.append(("g", lambda seg=seg: run(seg)))
            lo_i = hi_i
            if g_hi > g_lo:
                ph.append(("n", lambda a=g_lo, b=g_hi: works.append(
                    dist.all_reduce(fl.g[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))))
        ph.append(("n", lambda: ([w_.wait() for w_ in works], works.clear())))
        return ph

    def _buckets(self, plan):
        """[(end index into plan.bops, (flat lo, flat hi))]: parameters are laid out in registration order (stem,
        layer1, transition1, stage2, transition2, stage3, transition3, stage4, incre, head), the backward runs in the
        opposite order, so each stage boundary closes a contiguous tail of the flat gradient buffer."""
        if getattr(plan, "_bucket_cache", None) is not None:
            return plan._bucket_cache
        fl = self.flat
        first = {}
        for name, p, o, n in fl.entries:
            for st in ("stage2", "stage3", "stage4"):
                if ("." + st + ".") in ("." + name) and st not in first:
                    first[st] = o
        out, hi = [], fl.total
        for idx, st in plan.marks:                               # backward order: stage4, stage3, stage2
            if st in first:
                out.append((idx, (first[st], hi)))
                hi = first[st]
        out.append((len(plan.bops), (0, hi)))
        plan._bucket_cache = out
        return out

    def _optimizer(self, plan):
        fl = self.flat
        fl.step_count += 1
        a = fl.pack_args(self.lr, self.betas[0], self.betas[1], self.eps, self.wd, fl.step_count, 1.0, 1)
        rc = getattr(plan.lib, a[0])(*a[1:], torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError(f"bpb_adam_pack failed: {plan.lib.bpb_last_error().decode()}")
        self.model._param_version = getattr(self.model, "_param_version", 0) + 1   # cached eval plans re-pack lazily
        self._packed_version = self.model._param_version                           # (bpb_adam_pack just re-packed ours)

    def forward_backward(self, images, pids, masks=None):
        """Forward + loss + backward (gradients left in flat.g, summed over ranks); no parameter update."""
        plan = self._plan(images, masks)
        ver = getattr(self.model, "_param_version", 0)
        if ver != getattr(self, "_packed_version", None):       # load_state_dict / foreign update of the fp32 masters
            self.flat.pack_now()
            self._packed_version = ver
        plan.images.copy_(images, non_blocking=True)
        plan.pids.copy_(pids, non_blocking=True)
        if getattr(plan, "pids_rep", None) is not None:
            plan.pids_rep.copy_(pids.repeat(plan.K))
        if plan.masks_buf is not None:
            plan.masks_buf.copy_(masks, non_blocking=True)
        key = id(plan)
        if self.use_graph and self._world() == 1:
            if key not in self.graphs:
                self._launch(plan)                              # warm-up (first-call attribute setup)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(plan)
                self.graphs[key] = g
            self.graphs[key].replay()
        elif self.use_graph:
            # data parallel: one CUDA graph per kernel phase, the collectives run eagerly between the replays
            if key not in self.graphs:
                self._launch(plan)                              # warm-up: kernels AND communicators
                torch.cuda.synchronize()
                seq = []
                for kind, fn in self._phases(plan):
                    if kind == "g":
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            fn()
                        seq.append(g.replay)
                    else:
                        fn()                                    # keep every rank's collective sequence in step
                        seq.append(fn)
                torch.cuda.synchronize()
                self.graphs[key] = seq
            # (the collectives issued while capturing ran on stale buffers: the real step is always a full replay)
            for fn in self.graphs[key]:
                fn()
        else:
            self._launch(plan)                                  # (world > 1: gradient all-reduce inside, bucketed)
        return plan

    def step(self, images, pids, masks=None):
        plan = self.forward_backward(images, pids, masks)
        self._optimizer(plan)
        return plan.loss_out[0]

    __call__ = step

    def grads(self):
        """{parameter name: gradient tensor in the reference's shape} (conv weights as [cout, cin, k, k] views)."""
        fl = self.flat
        conv_w = {id(c.weight): c for _, _, c in fl.conv_ops.values()}
        out = {}
        for name, p, o, n in fl.entries:
            if id(p) not in fl.touched:
                continue
            out[name] = fl.conv_grad_view(conv_w[id(p)]) if id(p) in conv_w else fl.grad_view(p)
        return out


class _PublishGrads(torch.autograd.Function):
    """loss tensor whose .backward() publishes the gradients the fused step already computed (scaled by grad_output)
    into the parameters' .grad -- so `loss.backward(); optimizer.step()` of the reference engine
    (part_based_engine.py:101-102) keeps working with any torch optimizer."""

    @staticmethod
    def forward(ctx, loss_value, step, *params):
        ctx.step = step
        return loss_value.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        grads = ctx.step.grads()
        by_id = {id(p): g for (name, p, o, n), g in ((e, grads.get(e[0])) for e in ctx.step.flat.entries) if g is not None}
        outs = []
        for p in ctx.step._pub_params:
            g = by_id.get(id(p))
            outs.append(None if g is None else g * grad_out)
        return (None, None, *outs)


def autograd_loss(step: TrainStep, images, pids, masks=None, optimizer=None):
    """forward + loss + analytic backward now; returns a loss tensor that can be `.backward()`-ed like the reference's.
    With `optimizer` (a torch optimizer over model.parameters()) a post-step hook re-packs the bf16 operand copies."""
    plan = step.forward_backward(images, pids, masks)
    params = [p for _, p, _, _ in step.flat.entries if p.requires_grad]
    step._pub_params = params
    if optimizer is not None and not getattr(optimizer, "_bpb_hooked", False):
        def _bump(*_a, **_k):
            step.model._param_version = getattr(step.model, "_param_version", 0) + 1
        optimizer.register_step_post_hook(_bump)
        optimizer._bpb_hooked = True
    return _PublishGrads.apply(plan.loss_out[0], step, *params)
