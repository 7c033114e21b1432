#!/usr/bin/env python
"""bench.py -- images/s of the YOLO detection data path (letterbox -> Detect decode ->
confidence filter -> NMS) on BASELINE.json configs[1]: YOLOv5s, synthetic batch of 64
640x640 uint8 images per GPU, demo-mode NMS (conf 0.25, iou 0.45, best class, max_det 300).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N > 1); images shard across ranks with no collective on
the hot path ("weak" scaling: 64 images per GPU per step).  A step is one pass of the three
kernels over one batch.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

BATCH = 64
IMG = 640
CONF, IOU, MAX_DET = 0.25, 0.45, 300
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md


def baseline_metric():
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "images/sec pre+postproc at 640 (1/2/4/8 B200), HBM GB/s frac; vs host-CPU ref"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


WORKLOAD = ("configs[1]: YOLOv5s synthetic batch 64x640x640 letterbox+decode+NMS "
            "(conf 0.25, iou 0.45, best-class, max_det 300)")


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU is under load."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- reference arm
def cpu_reference_pass(imgs, levels):
    """The reference's CPU path on one sample: cv2 letterbox + normalise per image, torch
    Detect decode, torch + torchvision.ops.nms (oracle/ref_port.py, pinned to the live
    reference by tests/test_oracle_golden.py)."""
    from oracle import ref_port
    from tests import synth
    xs = [ref_port.preprocess(im, (IMG, IMG), is_bgr=True)[0] for im in imgs]
    pred, _ = ref_port.detect_decode(levels, synth.V5_ANCHORS, synth.STRIDES, "v5")
    dets = ref_port.nms(pred, conf_thres=CONF, iou_thres=IOU, max_det=MAX_DET)
    return xs, dets


def cpu_sample(n: int, seed: int = 0):
    from tests import synth
    imgs = list(synth.images_u8(n, IMG, IMG, seed=seed))
    levels = [torch.from_numpy(x) for x in synth.head_logits(n, seed=2 + seed, clusters=20)]
    return imgs, levels


def time_cpu(n_imgs: int, budget_s: float, max_passes: int = 1000):
    imgs, levels = cpu_sample(n_imgs)
    cpu_reference_pass(imgs, levels)                       # warm-up (first call is 10-20x slower)
    t0 = time.perf_counter()
    passes = 0
    while passes < max_passes:
        cpu_reference_pass(imgs, levels)
        passes += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return n_imgs * passes / dt, passes, dt


def run_reference(args, rank: int):
    if rank != 0:
        return
    n = 16
    imgs, levels = cpu_sample(n)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(imgs, levels)
    steps = max(1, min(args.steps, 40))                    # each step = one 16-image sample
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(imgs, levels)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    cores = torch.get_num_threads()
    sample = f"{n} images per step of the {BATCH}-image batch, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": baseline_metric(), "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "batch_per_step": n},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": sample, "os_cpu_count": os.cpu_count()},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true",
                    help="run the NMS on the main stream instead of overlapping it with the next batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from vision_kit_b200 import _lib
    from tests import synth
    from vision_kit_b200.pipeline import DetectPipeline
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: everything libraries print meanwhile (NCCL's version
    # banner, warnings) is routed to stderr by pointing fd 1 at fd 2 until the result is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic workload, one shard of BATCH images per rank
    imgs_host = torch.from_numpy(synth.images_u8(BATCH, IMG, IMG, seed=rank)).pin_memory()
    lv_host = [torch.from_numpy(x).pin_memory() for x in synth.head_logits(BATCH, seed=2 + rank, clusters=20)]
    imgs_dev = imgs_host.to(dev)
    lv_dev = [t.to(dev) for t in lv_host]
    pipe = DetectPipeline("v5", nc=80, img_sz=(IMG, IMG), batch=BATCH, conf_thres=CONF, iou_thres=IOU,
                          max_det=MAX_DET, swap_rb=True, device=dev, overlap=not args.no_overlap)
    pipe.plan_sources(list(imgs_dev))

    def step():
        pipe.preprocess()
        pipe.filter(lv_dev)
        pipe.nms()

    side = pipe.side if pipe.overlap else torch.cuda.current_stream()

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs (626 MB per step > 126 MB L2)
    K = args.steps
    # Events sit on the stream each kernel is launched on (NMS: the side stream when overlapping).  Per-kernel
    # events are recorded on every 4th step of the timed region (5 records cost ~20 us of host time, and
    # with 8 ranks on one host the loop must stay GPU-bound); the step time itself uses all K steps.
    EV = 4
    sampled = [k for k in range(K) if k % EV == 0]
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(5)] for k in sampled}
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n0 = _lib.launch_count()
    t_issue0 = time.perf_counter()
    t_begin.record()
    for k in range(K):
        if k % EV == 0:
            e = ev[k]
            e[0].record()
            pipe.preprocess()
            e[1].record()
            pipe.filter(lv_dev)
            e[2].record()
            e[3].record(side)       # queued behind the previous NMS on the side stream
            pipe.nms()
            e[4].record(side)
        else:
            pipe.preprocess()
            pipe.filter(lv_dev)
            pipe.nms()
    pipe.join()
    t_end.record()
    host_issue_us = (time.perf_counter() - t_issue0) / K * 1e6     # host time to enqueue one step (no waiting)
    barrier()
    launches = _lib.launch_count() - n0
    total_ms = t_begin.elapsed_time(t_end)
    kern_ms = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in sampled) / len(sampled) for i in range(2)]
    # NMS duration: from the later of (filter done, previous NMS done) to its own end
    nms_ms = 0.0
    for k in sampled:
        nms_ms += min(ev[k][2].elapsed_time(ev[k][4]), ev[k][3].elapsed_time(ev[k][4]))
    kern_ms.append(nms_ms / len(sampled))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * BATCH * K / (total_ms * 1e-3)
    if world > 1:                                       # launches of the whole job
        tl = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl.item())

    # ---- end to end through the public API with HOST buffers: H2D of the step's inputs
    # (uint8 images and the head's conv outputs), the three kernels, D2H of detections+counts
    Ke = max(1, min(args.e2e_steps, K))
    dets_host = torch.empty((BATCH, MAX_DET, 6), dtype=torch.float32).pin_memory()
    cnt_host = torch.empty((BATCH,), dtype=torch.int32).pin_memory()

    def e2e_step():
        imgs_dev.copy_(imgs_host, non_blocking=True)
        for d, h in zip(lv_dev, lv_host):
            d.copy_(h, non_blocking=True)
        pipe.preprocess()
        out = pipe.postprocess(lv_dev, join=True)
        dets_host.copy_(out.dets, non_blocking=True)
        cnt_host.copy_(out.counts, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Ke):
        e2e_step()
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * Ke / (float(te.item()) * 1e-3)
    h2d = imgs_host.numel() + sum(t_.numel() * 4 for t_ in lv_host)
    d2h = dets_host.numel() * 4 + cnt_host.numel() * 4
    clocks = sampler.result()
    n_dets = int(cnt_host.sum())
    n_cand = int(pipe.cand.counts.sum().item())

    # ---- roofline of the dominant kernel (algorithmic bytes)
    peak, peak_kind = measured_peak()
    src_bytes = BATCH * IMG * IMG * 3
    alg = {
        "lb_copy_kernel": src_bytes + BATCH * 3 * IMG * IMG * 4,
        "decode_filter_kernel": BATCH * pipe.rows * (pipe.cfg.nc + 5) * 4 + 8 * n_cand,
        "nms_staged_kernel": 24 * n_cand + BATCH * MAX_DET * 24,   # + the (idle) nms_image_kernel fallback launch
    }
    names = list(alg)
    kernels = {n_: {"ms": kern_ms[i], "algorithmic_bytes": alg[n_],
                    "achieved_gbs": alg[n_] / (kern_ms[i] * 1e-3) / 1e9,
                    "frac_of_peak": alg[n_] / (kern_ms[i] * 1e-3) / 1e9 / peak}
               for i, n_ in enumerate(names)}
    # the kernel that bounds a step: with the NMS overlapped on its side stream, the longest of the
    # two HBM-bound kernels on the main stream; otherwise the longest of all three
    crit = names[:2] if pipe.overlap else names
    dom = max(crit, key=lambda n_: kernels[n_]["ms"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac_of_peak"], "traffic": traffic,
                "peak_kind": f"of {peak_kind}",
                "note": ("achieved = algorithmic bytes (full conv-output read) / time; the kernel gathers only "
                         "surviving rows, so its DRAM traffic is below the algorithmic bytes and frac can exceed 1"
                         if dom == "decode_filter_kernel" else
                         ("duration measured while the NMS of the previous batch runs concurrently on the side stream "
                          "(it shares the SMs); the same kernel alone: 64 us = 0.94 of peak, profiles/r1z_kernels.txt"
                          if pipe.overlap else ""))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        v, passes, dt = time_cpu(8, args.cpu_seconds)
        cpu = {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"8 images of the batch x {passes} passes ({dt:.1f} s): cv2 letterbox+normalise, "
                         f"torch decode, torch+torchvision NMS", "os_cpu_count": os.cpu_count()}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps({
        "metric": baseline_metric(), "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                   "pipeline": "letterbox(u8 HWC->f32 NCHW /255) -> fused Detect decode+filter -> per-image sort+NMS",
                   "streams": ("NMS of batch k on a side stream, overlapped with letterbox+filter of batch k+1"
                               if pipe.overlap else "single stream"),
                   "parallelism": f"images sharded over {world} GPU(s), no collective on the hot path",
                   "l2": "inputs larger than L2: 627 MB read per step per GPU vs 126 MB L2, no flush needed",
                   "detections_per_step": n_dets, "candidates_per_step": n_cand,
                   "host_issue_us_per_step": round(host_issue_us, 1)},
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": Ke, "note": "PCIe-bound: the head's conv outputs (548 MB/step) are copied from host "
                                     "memory too, as the contract asks; in deployment they are produced on the GPU"},
        "gpu_launches": int(launches), "clocks": clocks,
    }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
