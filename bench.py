#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: images/sec, forward + GiLt loss, HRNet-W32 BPBreID K=5, batch 256 per GPU,
256x128 RGB, bf16 activations, synthetic data, random-init weights (num_classes 751).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|eval]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

One JSON line on rank 0.  `value` = device-timed whole-job images/s with the batch already resident in HBM;
`e2e` = the same step through the public API (bpbreid_b200.models / losses) starting from pinned HOST images with
the H2D copy and the D2H read of the loss inside the timed region.  `roofline` = the dominant kernel class timed
alone with CUDA events; `cpu_baseline` = the oracle port (CPU PyTorch fp32 restatement of the reference) on a
bounded sample of the same workload.  --impl reference times that CPU port for the driver's reference arm.
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

H, W, K_PARTS, NUM_CLASSES, PER_GPU_BATCH = 256, 128, 5, 751, 256
GFLOP_PER_IMAGE = 10.68            # conv+linear, 2*MAC, HRNet-W32 BPBreID K=5 @256x128


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


def make_cfg():
    from tests.golden_inputs import make_cfg as mk
    return mk("hrnet32", K_PARTS)


def synth_batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, H, W, generator=g)
    pids = (torch.arange(n) // 4).to(torch.int64)          # P x K sampler: 4 instances per identity
    return images, pids


# ---------------------------------------------------------------------------------------------- CPU port
def cpu_port_step_factory(batch):
    """The oracle port (oracle/model_oracle.py + oracle/loss_oracle.py): CPU fp32 restatement of the reference."""
    from oracle import loss_oracle as LO
    from oracle import model_oracle as MDL
    from bpbreid_b200 import models
    torch.manual_seed(0)
    m = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False, config=make_cfg())
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    images, pids = synth_batch(batch, 1)
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


def time_cpu_port(batch, steps, warmup):
    global _CPU_THREADS
    if _CPU_THREADS is None:
        _CPU_THREADS = pick_cpu_threads()
    torch.set_num_threads(_CPU_THREADS)
    step = cpu_port_step_factory(batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt


def run_reference_arm(args, rank):
    if rank != 0:
        return
    sample = 32
    ips, dt = time_cpu_port(sample, args.steps, min(args.warmup, 1))
    cores = _CPU_THREADS
    line = {"impl": "reference", "metric": "images/sec fwd+loss", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"HRNet-W32 BPBreID K=5 fwd+GiLt, 256x128, bounded sample of {sample} images/step "
                                   "of the bs256 batch (CPU)", "per_gpu_batch": PER_GPU_BATCH},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{sample}-image steps, fp32, oracle port of the reference (the Python reference "
                                       "cannot travel to the GPU box)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
KERNELS_PER_CALL = {"bpb_pixel_classify": 3, "bpb_triplet_fwd": 3, "bpb_ce_ls_fwd": 2}


def roofline_probe(dev, peaks):
    """Dominant kernel class = the tcgen05 3x3 conv at the stride-4 resolution (C=32, 64x32 map, batch 256): 150 of the
    662 launches and the largest single share of step time (profiles/launches_r1_step.md).  Its arithmetic intensity
    (9.66 GFLOP / 67 MB = 144 FLOP/B) is below the machine balance (~250 FLOP/B), so the bound is HBM.
    Timed with CUDA events around 8 back-to-back launches that rotate over 4 input/output pairs (4 x 67 MB > L2, so every
    launch reads cold data and the host-side launch gap is amortised)."""
    from bpbreid_b200 import _lib
    n, h, w, c = PER_GPU_BATCH, 64, 32, 32
    sets = []
    for _ in range(4):
        x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
        y = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        sets.append((x, y))
    wt = (torch.randn(c, 9 * c, device=dev) * 0.05).to(torch.bfloat16)
    tiles = _lib.query("bpb_conv_num_mtiles", n, h, w)
    stats = torch.empty(tiles, 2, c, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def call(i):
        x, y = sets[i % 4]
        _lib.call("bpb_conv_bf16", x.data_ptr(), wt.data_ptr(), y.data_ptr(), stats.data_ptr(), n, h, w, c, c, 3, 1, 0, s)

    for i in range(4):
        call(i)
    torch.cuda.synchronize()
    reps, per = 10, 8
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(per):
            call(i)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / per)
    ms = float(np.median(times))
    flops = 2.0 * n * h * w * c * 9 * c
    algo_bytes = n * h * w * c * 2 * 2 + 9 * c * c * 2 + tiles * 2 * c * 4
    achieved = algo_bytes / (ms * 1e-3) / 1e9
    traffic = None
    tf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "roofline_traffic_r1.json")
    if os.path.exists(tf):                                  # dram bytes per launch from one `ncu --set full` capture
        try:
            traffic = json.load(open(tf)).get("conv3x3_c32_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    return {"bound": "hbm", "kernel": "tc_gemm_kernel<32,bf16,raw+stats> conv3x3 32->32 @64x32 bs256",
            "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "ms_per_launch": ms, "algorithmic_bytes": algo_bytes,
            "tflops_at_this_rate": flops / (ms * 1e-3) / 1e12, "tensor_peak_tflops": peaks["bf16_tflops"],
            "peak_source": peaks["source"]}


def run_gpu_arm(args, rank, world):
    import torch.distributed as dist
    from bpbreid_b200 import distributed as D
    from bpbreid_b200 import losses, models
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    torch.manual_seed(0)
    model = models.build_model("bpbreid", num_classes=NUM_CLASSES, loss="part_based", pretrained=False,
                               use_gpu=True, config=make_cfg()).to(dev)
    model.train()
    model.use_cuda_graph = not args.no_graph
    model.materialize_spatial_features = False      # the 1920-ch map is an unused *output* of forward+loss
    crit = losses.GiLtLoss(None, use_visibility_scores=False, triplet_margin=0.3, use_gpu=True)
    n = PER_GPU_BATCH
    images_h, pids_h = synth_batch(n, 1 + rank)
    pids_h = (pids_h + rank * (n // 4)) % NUM_CLASSES
    images_pin, pids = images_h.pin_memory(), pids_h.to(dev)
    images_d = images_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        out = model(images_d)
        loss, _ = D.distributed_gilt(crit, out[0], out[1], out[2], pids)
        return loss

    # end-to-end: every step copies its 100 MB fp32 batch from pinned host memory and reads its loss back.  Like a
    # prefetching DataLoader, the H2D copy of step k+1 runs on a copy stream while step k computes (double-buffered).
    copy_stream = torch.cuda.Stream(dev)
    pins = [images_pin, images_h.clone().pin_memory()]
    stage = [torch.empty_like(images_d) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            stage[i % 2].copy_(pins[i % 2], non_blocking=True)
            ready[i % 2].record(copy_stream)

    def step_e2e(i):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        out = model(stage[i % 2])
        prefetch(i + 1)                                      # the other buffer's last reader (step i-1) has been synced
        loss, _ = D.distributed_gilt(crit, out[0], out[1], out[2], pids)
        return float(loss)                                   # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    plan = next(iter(model._plans.values()))
    if args.profile:
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile": "one step", "launches": sum(KERNELS_PER_CALL.get(op[2], 1) for op in plan.ops if op[0] is not None)}))
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()                                        # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = step_device(); e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    loss_val = float(loss)
    # end-to-end through the public API from pinned host memory
    prefetch(0)
    for i in range(2):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(2, 2 + args.steps):
        step_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = dev_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)
    launches_per_step = sum(KERNELS_PER_CALL.get(op[2], 1) for op in plan.ops if op[0] is not None) + 3 * 2 + 3 + 1 + 2
    roof = roofline_probe(dev, peaks)
    cpu_ips, cpu_dt = time_cpu_port(32, 2, 1) if not args.skip_cpu else (None, None)
    line = {
        "metric": "images/sec fwd+loss", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "HRNet-W32 BPBreID K=5, forward + GiLt loss (3 ID-CE + part-averaged batch-hard triplet), "
                               "train-mode BN, 256x128, num_classes 751, random-init weights",
                   "per_gpu_batch": n, "global_batch": n * world,
                   "parallelism": f"dp{world}: batch sharded, embeddings all-gathered for global-batch mining",
                   "l2": "working set >> 126 MB L2 and a 256 MB flush buffer is written between timed iterations",
                   "cuda_graph": bool(model.use_cuda_graph), "loss": loss_val,
                   "head": "evaluated on the 4 un-concatenated HRNet branch maps (linear identity); the [N,1920,64,32] "
                           "concat, which forward+loss never consumes, is not written"},
        "clocks": clocks,
        "e2e": {"value": world * n * args.steps / e2e_s, "unit": "images/s",
                "h2d_bytes_per_step": int(images_pin.numel() * 4), "d2h_bytes_per_step": 4,
                "pipeline": "double-buffered: the H2D copy of step k+1 (copy stream) overlaps the compute of step k; "
                            "every timed step still copies one batch and reads its loss"},
        "gpu_launches": int(launches_per_step * args.steps),
        "model_tflops": value * GFLOP_PER_IMAGE / 1e3,
        "model_frac_of_bf16_peak": value * GFLOP_PER_IMAGE / 1e3 / world / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
        "roofline": roof,
        "cpu_baseline": None if cpu_ips is None else {
            "value": cpu_ips, "unit": "images/s", "cores": _CPU_THREADS, "kind": "port",
            "sample": "2 steps of 32 images (same model/loss, fp32, oracle port of the reference on the host cores)"},
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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="run ONE forward+loss step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference for the CPU arm")
    run_gpu_arm(args, rank, world)


if __name__ == "__main__":
    main()
