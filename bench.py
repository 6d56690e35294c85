#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: images/sec, forward + GiLt loss, HRNet-W32 BPBreID K=5, batch 256 per GPU,
256x128 RGB, bf16 activations, synthetic data, random-init weights (num_classes 751).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload fwdloss|train|eval] [--height 256|384] [--batch B] [--scaling weak|strong]
                    [--no-extras] [--skip-cpu] [--skip-library]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

One JSON line on rank 0.
  value        device-timed whole-job images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e          the same step through the public API starting from pinned HOST images: H2D copy of every batch and the D2H
               read of the loss inside the timed region
  roofline     the dominant kernel class timed alone with CUDA events
  cpu_baseline the oracle port (CPU PyTorch fp32 restatement of the reference) on a bounded sample of the same workload
  extras       (default run only) further BASELINE configs measured in the same process, a few seconds each:
               train_step (fwd + loss + analytic backward + Adam), config2 (bs64 384x128), eval_config5 (4096 x 16384
               part distance + ranking, with its own HBM roofline), library_baseline (the reference port on ATen/cuDNN
               bf16 channels_last on the same GPU) and, for N > 1, strong scaling at global batch 256.
--workload train|eval make one of those the headline line instead; --impl reference times the CPU port for the driver's
reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_PARTS, NUM_CLASSES = 5, 751
GFLOP_PER_IMAGE = {256: 10.68, 384: 16.01}     # conv+linear, 2*MAC, HRNet-W32 BPBreID K=5


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.lines, self.gpu = None, [], gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = [s for s in sm if s > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


IBN_BACKBONE = "fastreid_resnet_ibn"


def make_cfg(backbone="hrnet32", parts=K_PARTS):
    from bpbreid_b200.config import make_cfg as mk
    return mk(backbone, parts)


def synth_batch(n, seed, h=256, w=128):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, h, w, generator=g)
    pids = (torch.arange(n) // 4).to(torch.int64)          # P x K sampler: 4 instances per identity
    return images, pids


# ---------------------------------------------------------------------------------------------- CPU port
def cpu_port_step_factory(batch, h=256, w=128):
    """The oracle port (oracle/model_oracle.py + oracle/loss_oracle.py): CPU fp32 restatement of the reference."""
    from oracle import loss_oracle as LO
    from oracle import model_oracle as MDL
    from bpbreid_b200 import models
    torch.manual_seed(0)
    m = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False, config=make_cfg())
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, pids = synth_batch(batch, 1, h, w)
    pids_np = pids.numpy() % NUM_CLASSES

    def step():
        with torch.no_grad():
            emb, vis, ids, *_ = MDL.bpbreid_forward(sd, images, "hrnet32", K_PARTS, training=True)
        loss, _ = LO.gilt({k: v.numpy() for k, v in emb.items()}, {k: v.numpy() for k, v in vis.items()},
                          {k: v.numpy() for k, v in ids.items()}, pids_np)
        return float(loss)
    return step


def pick_cpu_threads():
    """oneDNN convolutions on the tiny HRNet maps slow down badly when oversubscribed (128 threads: 0.5 img/s, measured
    on the GPU box): calibrate once on an 8-image step and keep the fastest thread count."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    if len(cands) == 1:
        return cands[0]
    step = cpu_port_step_factory(8)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    return best


_CPU_THREADS = None


def time_cpu_port(batch, steps, warmup, h=256, w=128):
    global _CPU_THREADS
    if _CPU_THREADS is None:
        _CPU_THREADS = pick_cpu_threads()
    torch.set_num_threads(_CPU_THREADS)
    step = cpu_port_step_factory(batch, h, w)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def workload_config(args, world):
    """`config` of the JSON line -- identical for both arms (the reference arm runs a bounded sample of THIS workload)."""
    h = args.height
    per_gpu = args.batch if args.scaling == "weak" else args.batch // world
    what = {"fwdloss": "forward + GiLt loss (3 ID-CE + part-averaged batch-hard triplet)",
            "train": "forward + GiLt loss + analytic backward + Adam step"}[args.workload]
    return {"workload": f"HRNet-W32 BPBreID K=5, {what}, train-mode BN, {h}x128, num_classes 751, random-init weights",
            "per_gpu_batch": per_gpu, "global_batch": per_gpu * world,
            "parallelism": f"dp{world}: batch sharded; one all-gather of the part embeddings, every rank mines its own "
                           "anchor rows of the global pairwise matrix"
                           + ("; bucketed NCCL all-reduce of the flat gradient" if args.workload == "train" else ""),
            "l2": "working set >> 126 MB L2 and a 256 MB flush buffer is written between timed iterations",
            "cuda_graph": True,
            "head": "evaluated on the 4 un-concatenated HRNet branch maps (linear identity); the [N,1920,H/4,W/4] "
                    "concat, which forward+loss never consumes, is not written"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    if args.workload != "fwdloss":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference arm times forward + GiLt loss only"}), flush=True)
        return
    sample = 32
    warm = max(min(args.warmup, 3), 1)
    ips, dt = time_cpu_port(sample, args.steps, warm, args.height, 128)
    cores = _CPU_THREADS
    line = {"impl": "reference", "metric": "images/sec fwd+loss", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(args.gpus, 1)),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"every step is a {sample}-image sample of the configured batch (same model / loss, fp32, "
                                       "host cores; the 'l2' / 'cuda_graph' / 'head' notes of config describe the GPU arm), "
                                       "oracle port of the reference (the Python reference cannot travel to the GPU box)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
KERNELS_PER_CALL = {"bpb_pixel_classify": 3, "bpb_triplet_fwd": 3, "bpb_ce_ls_fwd": 2, "bpb_triplet_fwd_rows": 3}


def _cuda_time(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def roofline_probe(dev, peaks, batch):
    """Dominant kernel class = the tcgen05 3x3 conv at the stride-4 resolution (C=32, 64x32 map): 150 launches per step
    and the largest single share of step time (profiles/launches_r2_step.md).  Its arithmetic intensity
    (9.66 GFLOP / 67 MB = 144 FLOP/B at batch 256) is below the machine balance (~250 FLOP/B), so the bound is HBM.
    Timed with CUDA events (the launches go to torch's current stream) around 8 back-to-back launches that rotate over 4
    input/output pairs (4 x 67 MB > L2, so every launch reads cold data and the host-side launch gap is amortised)."""
    from bpbreid_b200 import _lib
    n, h, w, c = batch, 64, 32, 32
    sets = []
    for _ in range(4):
        x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
        y = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        sets.append((x, y))
    wt = (torch.randn(c, 9 * c, device=dev) * 0.05).to(torch.bfloat16)
    tiles = _lib.query("bpb_conv_num_mtiles", n, h, w)
    stats = torch.empty(tiles, 2, c, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    it = [0]

    def call():
        x, y = sets[it[0] % 4]
        it[0] += 1
        _lib.call("bpb_conv_bf16", x.data_ptr(), wt.data_ptr(), y.data_ptr(), stats.data_ptr(), n, h, w, c, c, 3, 1, 0, s)

    for _ in range(4):
        call()
    torch.cuda.synchronize()
    ms = float(np.median([_cuda_time(call, 8) for _ in range(10)]))
    flops = 2.0 * n * h * w * c * 9 * c
    algo_bytes = n * h * w * c * 2 * 2 + 9 * c * c * 2 + tiles * 2 * c * 4
    achieved = algo_bytes / (ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    tf = os.path.join(ROOT, "profiles", "roofline_traffic_r2.json")
    if os.path.exists(tf):            # dram bytes per launch of THIS kernel from this round's `ncu --set full` capture
        try:
            d = json.load(open(tf))
            traffic, tsrc = d.get("conv3x3_c32_bs256_dram_bytes_per_launch"), d.get("source")
        except (OSError, ValueError):
            traffic = None
    if batch != 256:
        traffic = None
    return {"bound": "hbm", "kernel": f"tc_gemm_kernel<32,bf16,raw+stats> conv3x3 32->32 @64x32 bs{n}",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "traffic_source": tsrc, "ms_per_launch": ms, "timing": "CUDA events, launch stream",
            "algorithmic_bytes": algo_bytes, "tflops_at_this_rate": flops / (ms * 1e-3) / 1e12,
            "tensor_peak_tflops": peaks["bf16_tflops"], "peak_source": peaks["source"]}


def eval_config5(dev, peaks, nq=4096, ng=16384, p=6, d=512, reps=5):
    """BASELINE config 5: K=5 (+foreground) part distance 4096 x 16384 + market1501 ranking, inputs resident in HBM."""
    from bpbreid_b200 import _lib, metrics
    rs = np.random.RandomState(0)
    qf = torch.nn.functional.normalize(torch.from_numpy(rs.randn(nq, p, d).astype(np.float32)), dim=-1).to(dev)
    gf = torch.nn.functional.normalize(torch.from_numpy(rs.randn(ng, p, d).astype(np.float32)), dim=-1).to(dev)
    qv, gv = torch.from_numpy(rs.rand(nq, p) > 0.15).to(dev), torch.from_numpy(rs.rand(ng, p) > 0.15).to(dev)
    q_pids, g_pids = rs.randint(0, 1500, nq), rs.randint(0, 1500, ng)
    q_cam, g_cam = rs.randint(0, 6, nq), rs.randint(0, 6, ng)
    dist = torch.empty(nq, ng, device=dev)
    part = torch.empty(p, nq, ng, device=dev)
    scratch = torch.empty(_lib.query("bpb_part_distance_scratch_floats", nq, ng, p), device=dev)
    out = (dist, part, scratch)

    def run_dist():
        metrics.part_distance_device(qf, gf, qv, gv, "mean", "euclidean", 0, True, out)
    run_dist()
    torch.cuda.synchronize()
    d_ms = float(np.median([_cuda_time(run_dist, 1) for _ in range(reps)]))
    ids = [torch.as_tensor(a, dtype=torch.int64).to(dev) for a in (q_pids, g_pids, q_cam, g_cam)]

    def run_rank():
        return metrics.rank_device(dist, *ids, 50)
    res = run_rank()
    torch.cuda.synchronize()
    r_ms = float(np.median([_cuda_time(run_rank, 1) for _ in range(reps)]))
    algo = (nq + ng) * p * d * 4 + nq * ng * 4 * (p + 1)
    # optional k-reciprocal re-ranking (part_based_engine.py:218-226): q-q and g-g distances + rerank.py on the device
    rr = None
    try:
        from bpbreid_b200 import rerank

        def run_rerank():
            qq, _ = metrics.part_distance_device(qf, qf, qv, qv, "mean", "euclidean", 0, False)
            gg, _ = metrics.part_distance_device(gf, gf, gv, gv, "mean", "euclidean", 0, False)
            return rerank.re_ranking_device(dist, qq, gg)
        run_rerank()
        torch.cuda.synchronize()
        rr = {"ms": float(np.median([_cuda_time(run_rerank, 1) for _ in range(3)])),
              "what": "q-q + g-g part distances and k-reciprocal re-ranking (k1=20, k2=6), device-resident"}
    except Exception as e:
        rr = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    return {"rerank": rr, "workload": f"part distance {nq}x{ng}, P={p}, D={d}, bool visibility, mean + market1501 CMC/mAP",
            "distance_ms": d_ms, "rank_ms": r_ms, "mAP": float(res["mAP"].item()),
            "roofline": {"bound": "hbm", "kernel": "part_distance_tc_kernel (+ norms, masks, fix-up)",
                         "algorithmic_bytes": algo, "achieved": algo / (d_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": algo / (d_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "pairs_per_s": nq * ng / ((d_ms + r_ms) * 1e-3)}


def eval_end_to_end_leg(dev, nq=3368, ng=15913, batch=256):
    """The whole evaluation (part_based_engine.py:168-295) at Market-1501 scale through the engine mirror: feature extraction
    of nq + ng images (H2D from pinned host memory every batch), L2 normalisation, part distance, k-reciprocal re-ranking
    (q-q, g-g distances included), CMC / mAP -- everything after the copy stays on the device.  Wall clock around a sync."""
    from bpbreid_b200 import engine, models
    torch.manual_seed(0)
    cfg = make_cfg()
    m = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False, use_gpu=True,
                           config=cfg).to(dev).eval()
    m.use_cuda_graph = True
    m.materialize_spatial_features = False
    eng = engine.ImagePartBasedEngine(model=m, config=cfg, mask_filtering_testing=True)
    host = synth_batch(batch, 5, 256, 128)[0].pin_memory()
    rs = np.random.RandomState(0)

    def loader(n, cams):
        for lo in range(0, n, batch):
            k = min(batch, n - lo)
            yield {"image": host[:k], "pid": rs.randint(0, 751, k), "camid": rs.randint(*cams, k)}
    out = {}
    for rerank in (False, True):
        for _ in range(2):                                          # first pass builds the plans / graphs
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            cmc, m_ap = eng._evaluate(loader(nq, (0, 3)), loader(ng, (3, 6)), normalize_feature=True, rerank=rerank)
            torch.cuda.synchronize()
            sec = time.perf_counter() - t0
        out["with_rerank_s" if rerank else "seconds"] = sec
    out.update({"images": nq + ng, "images_per_s": (nq + ng) / out["seconds"],
                "what": f"{nq} query + {ng} gallery images: extraction (eval mode, fused epilogues, H2D per batch) + "
                        "normalise + part distance + CMC/mAP, device-resident; second figure adds k-reciprocal re-ranking"})
    return out


def input_pipeline_leg(dev, batch=256):
    """decoded Market-1501-sized samples (RGB uint8 128x64 + 36 pifpaf fields 16x8) -> model inputs at
    384x128 (every shipped YAML's size) through bpbreid_b200.data.DeviceInputPipeline: host packing into pinned memory,
    one H2D copy per stream of samples, two kernel launches; timed end to end on the host clock around a device sync."""
    from bpbreid_b200 import data
    rs = np.random.RandomState(0)
    images = [rs.randint(0, 256, (128, 64, 3)).astype(np.uint8) for _ in range(batch)]
    fields = [rs.rand(36, 16, 8).astype(np.float32) for _ in range(batch)]
    pipe = data.DeviceInputPipeline(384, 128, transforms=["rc", "re"], masks_preprocess="five_v", device=dev)
    params = pipe.draw_params(batch, np.random.RandomState(1))
    for _ in range(2):
        pipe(images, fields, params)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        pipe(images, fields, params)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ms = float(np.median(ts)) * 1e3
    return {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_batch": ms, "batch": batch,
            "what": "uint8 128x64 + 36x16x8 fields -> [N,3,384,128] f32 + [N,6,96,32] masks; host pack + H2D + 2 kernels "
                    "(wall clock); cv2-exact resize"}


def library_baseline(dev, batch, h):
    """The reference port on ATen / cuDNN (bf16 autocast, channels_last) on the same GPU: forward + GiLt, same batch.
    This is what the reference would run on this machine (it ships no kernels of its own)."""
    from oracle import model_oracle as MDL
    from oracle import loss_oracle_torch as LT
    from bpbreid_b200 import models
    torch.manual_seed(0)
    m = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False, config=make_cfg())
    sd = {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()}
    for k, v in sd.items():
        if v.dim() == 4:
            sd[k] = v.contiguous(memory_format=torch.channels_last)
    images, pids = synth_batch(batch, 1, h, 128)
    images = images.to(dev).contiguous(memory_format=torch.channels_last)
    pids = (pids % NUM_CLASSES).to(dev)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            emb, vis, ids, *_ = MDL.bpbreid_forward(sd, images, "hrnet32", K_PARTS, training=True)
            return LT.gilt({k: v.float() for k, v in emb.items()}, {k: v.float() for k, v in ids.items()}, pids)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = _cuda_time(step, 3)
    out = {"value": batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
           "kind": "reference port (oracle/model_oracle.py) on ATen/cuDNN, bf16 autocast, channels_last, eager"}
    # the same through autograd + fused Adam: what the reference's training step (part_based_engine.py:98-103) costs here
    try:
        leaves = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
        opt = torch.optim.Adam(leaves, lr=3.5e-4, weight_decay=5e-4, fused=True)

        def train_step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                emb, vis, ids, *_ = MDL.bpbreid_forward(sd, images, "hrnet32", K_PARTS, training=True)
                loss = LT.gilt({k: v.float() for k, v in emb.items()}, {k: v.float() for k, v in ids.items()}, pids)
            loss.backward()
            opt.step()
        for _ in range(2):
            train_step()
        torch.cuda.synchronize()
        tms = _cuda_time(train_step, 3)
        out["train"] = {"value": batch / (tms * 1e-3), "unit": "images/s", "ms_per_step": tms,
                        "what": "forward + GiLt + torch autograd backward + fused Adam, same port, bf16 autocast"}
    except Exception as e:                                         # the forward-only leg stands on its own
        out["train"] = {"error": repr(e)[:300]}
    return out


class FwdLoss:
    """forward + GiLt loss through the public API (models.build_model / losses.GiLtLoss / distributed_gilt)."""

    def __init__(self, dev, batch, h, rank, world, use_graph=True, backbone="hrnet32", parts=K_PARTS):
        from bpbreid_b200 import distributed as D
        from bpbreid_b200 import losses, models
        torch.manual_seed(0)
        self.D, self.world = D, world
        self.model = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False,
                                        use_gpu=True, config=make_cfg(backbone, parts)).to(dev)
        self.model.train()
        self.model.use_cuda_graph = use_graph
        self.model.materialize_spatial_features = False      # the 1920-ch map is an unused *output* of forward+loss
        self.crit = losses.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, use_gpu=True)
        self.cache = {}
        images_h, pids_h = synth_batch(batch, 1 + rank, h, 128)
        self.pids = ((pids_h + rank * (batch // 4)) % NUM_CLASSES).to(dev)
        self.images_h = images_h
        self.images_d = images_h.to(dev)
        self.dev, self.batch = dev, batch

    def step(self, images=None):
        out = self.model(self.images_d if images is None else images)
        loss, _ = self.D.distributed_gilt(self.crit, out[0], out[1], out[2], self.pids, cache=self.cache)
        return loss

    def launches(self):
        plan = next(iter(self.model._plans.values()))
        return sum(KERNELS_PER_CALL.get(op[2], 1) for op in plan.ops if op[0] is not None) + 3 * 2 + 3 + 1 + 2


class TrainRun:
    """fwd + loss + analytic backward + Adam (bpbreid_b200.train.TrainStep)."""

    def __init__(self, dev, batch, h, rank, world, use_graph=True):
        from bpbreid_b200 import losses, models, train
        torch.manual_seed(0)
        self.model = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False,
                                        use_gpu=True, config=make_cfg()).to(dev)
        self.model.materialize_spatial_features = False
        crit = losses.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, use_gpu=True)
        self.ts = train.TrainStep(self.model, crit, None, use_graph=use_graph)
        images_h, pids_h = synth_batch(batch, 1 + rank, h, 128)
        self.pids = ((pids_h + rank * (batch // 4)) % NUM_CLASSES).to(dev)
        self.images_h = images_h
        self.images_d = images_h.to(dev)
        self.batch = batch

    def step(self, images=None):
        return self.ts.step(self.images_d if images is None else images, self.pids)

    def launches(self):
        plan = next(iter(self.ts.plans.values()))
        fwd = sum(KERNELS_PER_CALL.get(op[2], 1) for op in plan.ops if op[0] is not None)
        return fwd + len(plan.loss_ops) + sum(1 for op in plan.bops if op[0] is not None) + 3


def timed_region(runner, args, dev, world, flush):
    """W warm-up steps, then K timed steps (CUDA events around each step on the launch stream, L2 flushed between)."""
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        runner.step()
    barrier()
    evs = []
    for _ in range(args.steps):
        flush.zero_()                                        # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = runner.step(); e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / args.steps, float(loss)


def e2e_region(runner, args, dev, world):
    """every step copies its fp32 batch from pinned host memory and reads its loss back.  Like a prefetching DataLoader,
    the H2D copy of step k+1 runs on a copy stream while step k computes (double-buffered)."""
    import torch.distributed as dist
    copy_stream = torch.cuda.Stream(dev)
    pins = [runner.images_h.pin_memory(), runner.images_h.clone().pin_memory()]
    stage = [torch.empty_like(runner.images_d) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            stage[i % 2].copy_(pins[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def step(i):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        loss = runner.step(stage[i % 2])
        prefetch(i + 1)                                      # the other buffer's last reader (step i-1) has been synced
        return float(loss)                                   # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    prefetch(0)
    for i in range(2):
        step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(2, 2 + args.steps):
        step(i)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), int(pins[0].numel() * 4)


def quick(runner, dev, world, steps=5, warmup=3):
    """short device-timed measurement for the extras"""
    a = argparse.Namespace(steps=steps, warmup=warmup)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms, loss = timed_region(runner, a, dev, world, flush)
    return ms, loss


def run_gpu_arm(args, rank, world):
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    h = args.height
    per_gpu = args.batch if args.scaling == "weak" else args.batch // world
    if args.scaling == "strong" and args.batch % (4 * world):
        raise SystemExit("strong scaling: the global batch must split into whole identity groups per rank")

    if args.workload == "eval":
        if rank == 0:
            ev = eval_config5(dev, peaks)
            line = {"metric": "query-gallery pairs/sec part-distance + rank (config 5)", "value": ev["pairs_per_s"],
                    "unit": "pairs/s", "n_gpus": 1, "steps": 5, "warmup": 1, "ms_per_step": ev["distance_ms"] + ev["rank_ms"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
                    "config": {"workload": ev["workload"]}, "roofline": ev["roofline"], "eval": ev, "gpu_launches": 8 * 5}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    runner = (TrainRun if args.workload == "train" else FwdLoss)(dev, per_gpu, h, rank, world, not args.no_graph)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    if args.profile:
        for _ in range(3):
            runner.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        runner.step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile": "one step", "launches": runner.launches()}))
        return
    for _ in range(max(args.warmup, 3)):                     # plans, graphs, NCCL communicators
        runner.step()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ms_per_step, loss_val = timed_region(runner, args, dev, world, flush)
    e2e_s, h2d = e2e_region(runner, args, dev, world)
    clocks = sampler.stop() if rank == 0 else None
    launches = runner.launches()

    extras = {}
    if not args.no_extras and args.workload == "fwdloss" and args.scaling == "weak":
        del runner
        torch.cuda.empty_cache()
        try:
            tr = TrainRun(dev, per_gpu, h, rank, world)
            ms, l = quick(tr, dev, world)
            extras["train_step"] = {"value": world * per_gpu / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
                                    "what": "fwd + GiLt + analytic backward (tcgen05 wgrad/dgrad) + Adam"
                                            + (", bucketed NCCL gradient all-reduce" if world > 1 else ", one CUDA graph"),
                                    "loss": l, "gpu_launches_per_step": tr.launches()}
            del tr
        except Exception as e:                               # an extra must never take the headline down with it
            extras["train_step"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        if world > 1 and args.batch % (4 * world) == 0:
            try:
                st = FwdLoss(dev, args.batch // world, h, rank, world)
                ms, l = quick(st, dev, world)
                extras["strong_scaling_global256"] = {"value": args.batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
                                                      "per_gpu_batch": args.batch // world,
                                                      "what": "SURVEY 8(e) config 3: global batch 256 sharded over the ranks"}
                del st
            except Exception as e:
                extras["strong_scaling_global256"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        if world == 1 and rank == 0:
            try:
                c2 = FwdLoss(dev, 64, 384, rank, world)
                ms, l = quick(c2, dev, world)
                extras["config2_bs64_384x128"] = {"value": 64 / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
                                                  "model_tflops": 64 / (ms * 1e-3) * GFLOP_PER_IMAGE[384] / 1e3, "loss": l}
                del c2
            except Exception as e:
                extras["config2_bs64_384x128"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            try:                                             # BASELINE config 4's per-GPU share (512 / 8), in bf16
                c4 = FwdLoss(dev, 64, 256, rank, world, backbone=IBN_BACKBONE, parts=8)
                ms, l = quick(c4, dev, world)
                extras["config4_ibn_k8_bs64_per_gpu"] = {
                    "value": 64 / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "loss": l,
                    "what": "fast-reid ResNet-50-IBN, K=8, 256x128, forward + GiLt, batch 64 = one GPU's share of config 4 "
                            "(bs512 on 8 GPUs); bf16 activations (the reference config says fp16: not built)"}
                del c4
            except Exception as e:
                extras["config4_ibn_k8_bs64_per_gpu"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            try:
                from bpbreid_b200 import models
                torch.manual_seed(0)
                me = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False,
                                        use_gpu=True, config=make_cfg()).to(dev).eval()
                me.use_cuda_graph = True
                me.materialize_spatial_features = False
                xe = synth_batch(per_gpu, 3, h, 128)[0].to(dev)
                with torch.no_grad():
                    for _ in range(3):
                        me(xe)
                    torch.cuda.synchronize()
                    ms = _cuda_time(lambda: me(xe), 5)
                pl = next(iter(me._plans.values()))
                extras["eval_forward"] = {"value": per_gpu / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
                                          "gpu_launches_per_step": sum(1 for op in pl.ops if op[0] is not None),
                                          "what": "feature extraction (eval mode): BatchNorm / residual / ReLU in the conv epilogue"}
                del me, xe
            except Exception as e:
                extras["eval_forward"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            try:
                extras["eval_end_to_end_market_scale"] = eval_end_to_end_leg(dev)
            except Exception as e:
                extras["eval_end_to_end_market_scale"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            try:
                extras["input_pipeline"] = input_pipeline_leg(dev)
            except Exception as e:
                extras["input_pipeline"] = {"error": repr(e)[:300]}
            try:
                extras["eval_config5"] = eval_config5(dev, peaks)
            except Exception as e:
                extras["eval_config5"] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
            if not args.skip_library:
                try:
                    extras["library_baseline"] = library_baseline(dev, per_gpu, h)
                except Exception as e:
                    extras["library_baseline"] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * per_gpu / (ms_per_step * 1e-3)
    roof = roofline_probe(dev, peaks, per_gpu)
    cpu_ips, cpu_dt = time_cpu_port(32, 2, 1, h, 128) if not args.skip_cpu else (None, None)
    gf = GFLOP_PER_IMAGE.get(h, 10.68) * (3.0 if args.workload == "train" else 1.0)
    line = {
        "metric": "images/sec fwd+loss" if args.workload == "fwdloss" else "images/sec fwd+bwd+opt", "value": value,
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world), "loss": loss_val,
        "clocks": clocks,
        "e2e": {"value": world * per_gpu * args.steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "pipeline": "double-buffered: the H2D copy of step k+1 (copy stream) overlaps the compute of step k; "
                            "every timed step still copies one batch and reads its loss"},
        "gpu_launches": int(launches * args.steps),
        "model_tflops": value * gf / 1e3,
        "model_frac_of_bf16_peak": value * gf / 1e3 / world / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
        "roofline": roof,
        "cpu_baseline": None if cpu_ips is None else {
            "value": cpu_ips, "unit": "images/s", "cores": _CPU_THREADS, "kind": "port",
            "sample": "2 steps of 32 images (same model/loss, fp32, oracle port of the reference on the host cores)"},
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fwdloss", choices=["fwdloss", "train", "eval"])
    ap.add_argument("--height", type=int, default=256, choices=[256, 384])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (weak scaling) or global batch (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly (kernel-by-kernel profiling)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-library", action="store_true")
    ap.add_argument("--lib", default=None, help="A/B aid: load this build of libbpbreid_sm100.so instead of the in-tree one")
    ap.add_argument("--profile", action="store_true",
                    help="run ONE step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE JSON line: anything a library writes to fd 1 (NCCL prints its version banner there)
    # goes to stderr instead, and the line is written through a private copy of the original descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.lib:
        from bpbreid_b200 import _lib
        _lib.LIB_PATH = os.path.abspath(args.lib)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    run_gpu_arm(args, rank, world)


if __name__ == "__main__":
    main()
