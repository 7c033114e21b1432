#!/usr/bin/env python
"""Per-kernel roofline table on the GPU box (CUDA events, warm, inputs > L2 where possible):
letterbox (identity / mixed-size resize, f32 / bf16), Detect decode (materialised, +raw),
filter_pred, fused decode_filter and NMS in demo and eval mode.

    python profiles/kernel_bench.py [--batch 64] > gpurun_out/kernels.json
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vision_kit_b200 import _lib, ops
from tests import synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
B = args.batch
dev = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
res = {}


def timeit(fn, iters=args.iters, warm=5):
    """Mean device time of one call: the calls are captured into a CUDA graph (the Python around a C call
    costs tens of microseconds, more than the short kernels take) and the graph is replayed."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
    except Exception as e:                                   # not capturable: time the eager loop
        print(f"(graph capture failed: {e}; eager timing)", file=sys.stderr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    res[name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 2), "GBps": round(gbs, 1),
                 "frac_of_measured_peak": round(gbs / PEAK, 3), "images_per_s": round(B / (ms * 1e-3)), "note": note}
    print(f"{name:34s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.3f} of peak  {B/(ms*1e-3):12.0f} img/s  {note}", file=sys.stderr)


# ---- letterbox
imgs = torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev)
for dt, nm, sz in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
    plan = ops.LetterboxPlan(list(imgs), (640, 640))
    out = torch.empty((B, 3, 640, 640), dtype=dt, device=dev)
    ms = timeit(lambda: plan.run(out, swap_rb=True))
    report(f"letterbox identity 640 {nm}", ms, B * (640 * 640 * 3 + 3 * 640 * 640 * sz))
sizes = synth.mixed_sizes(B, seed=5)
srcs = [torch.from_numpy(synth.image_u8(h, w, 50 + i)).to(dev) for i, (h, w) in enumerate(sizes)]
src_bytes = sum(h * w * 3 for h, w in sizes)
for dt, nm, sz in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
    plan = ops.LetterboxPlan(srcs, (640, 640))
    out = torch.empty((B, 3, 640, 640), dtype=dt, device=dev)
    ms = timeit(lambda: plan.run(out, swap_rb=True))
    report(f"letterbox mixed 480-1280 {nm}", ms, src_bytes + B * 3 * 640 * 640 * sz, "config 5 sources")
# pure up-scale and exact 2x
for (h, w), nm in (((480, 480), "up 480"), ((1280, 1280), "down 2x 1280")):
    s2 = [torch.from_numpy(synth.image_u8(h, w, 7)).to(dev) for _ in range(B)]
    plan = ops.LetterboxPlan(s2, (640, 640))
    out = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
    ms = timeit(lambda: plan.run(out, swap_rb=True))
    report(f"letterbox {nm} f32", ms, B * (h * w * 3 + 3 * 640 * 640 * 4))
del srcs, s2

# ---- decode / filter / nms
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
in_bytes = B * 25200 * 85 * 4
ms = timeit(lambda: ops.detect_decode(cfg, lv))
report("detect_decode -> pred", ms, 2 * in_bytes)
ms = timeit(lambda: ops.detect_decode(cfg, lv, want_raw=True))
report("detect_decode -> pred + raw", ms, 3 * in_bytes, "the reference's full return value")
pred = ops.detect_decode(cfg, lv)
lv16 = [t.half() for t in lv]
ms = timeit(lambda: ops.detect_decode(cfg, lv16))
report("detect_decode fp16 in -> pred", ms, in_bytes // 2 + in_bytes, "AMP eval: fp16 conv outputs")
for mode, conf, ml, iou in (("demo", 0.25, False, 0.45), ("eval", 0.001, True, 0.6)):
    buf = ops.filter_pred(pred, conf, ml)
    ncand = int(buf.counts.sum())
    ms = timeit(lambda: ops.filter_pred(pred, conf, ml, buf=buf))
    report(f"filter_pred {mode}", ms, in_bytes + 8 * ncand, f"{ncand // B} candidates/img; bytes = full pred read")
    buf2 = ops.decode_filter(cfg, lv, conf, ml)
    ms = timeit(lambda: ops.decode_filter(cfg, lv, conf, ml, buf=buf2))
    report(f"decode_filter {mode}", ms, in_bytes + 8 * ncand)
    buf16 = ops.decode_filter(cfg, lv16, conf, ml)
    ms = timeit(lambda: ops.decode_filter(cfg, lv16, conf, ml, buf=buf16))
    report(f"decode_filter {mode} fp16 in", ms, in_bytes // 2 + 8 * int(buf16.counts.sum()), "bytes = full fp16 conv-output read")
    outb = ops.nms_batched(buf2, iou)
    ms = timeit(lambda: ops.nms_batched(buf2, iou, out=outb), iters=10)
    # bytes the NMS has to move: every candidate key once (8 B), key + box (24 B) of the candidates it orders and
    # examines (all of them at demo thresholds, at most the top list at eval thresholds), the detections
    nms_bytes = 8 * ncand + 24 * min(ncand, B * ops.LIST_CAP) + B * 300 * 24
    report(f"nms {mode}", ms, nms_bytes, f"{int(outb.counts.sum()) // B} dets/img; latency-bound (one CTA per image after the select pass)")
    if mode == "eval":
        outa = ops.nms_batched(buf2, iou, agnostic=True)
        ms = timeit(lambda: ops.nms_batched(buf2, iou, agnostic=True, out=outa), iters=10)
        report("nms eval agnostic", ms, nms_bytes, f"{int(outa.counts.sum()) // B} dets/img")
        h = B // 2
        bufh = ops.decode_filter(cfg, [t[:h] for t in lv], conf, ml)
        outh = ops.nms_batched(bufh, iou)
        ms = timeit(lambda: ops.nms_batched(bufh, iou, out=outh), iters=10)
        report(f"nms eval, {h} images", ms * 2, nms_bytes, f"(ms, GB/s scaled to {B} images) actual {ms*1e3:.1f} us per {h}")
# plain logits (no planted clusters): SURVEY.md §8d base workload
lv0 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=0)]
buf3 = ops.decode_filter(cfg, lv0, 0.25, False)
ms = timeit(lambda: ops.decode_filter(cfg, lv0, 0.25, False, buf=buf3))
report("decode_filter demo (no clusters)", ms, in_bytes + 8 * int(buf3.counts.sum()), f"{int(buf3.counts.sum()) // B} candidates/img")
# ---- evaluator matching after eval-mode NMS, and the eval loader's ingest
buf4 = ops.decode_filter(cfg, lv, 0.001, True)
nout = ops.nms_batched(buf4, 0.6)
rng = np.random.Generator(np.random.PCG64(3))
dets_h, cnt_h = nout.dets.cpu().numpy(), nout.counts.cpu().numpy()
labs, offs, shp = [], [0], []
for b in range(B):
    k = int(cnt_h[b])
    pick = rng.choice(k, size=min(k, 25), replace=False)
    bx = dets_h[b, pick, :4]
    labs.append(np.concatenate([np.full((len(pick), 1), b, np.float32), dets_h[b, pick, 5:6],
                                np.stack([(bx[:, 0] + bx[:, 2]) / 2, (bx[:, 1] + bx[:, 3]) / 2, bx[:, 2] - bx[:, 0], bx[:, 3] - bx[:, 1]], 1)], 1))
    offs.append(offs[-1] + len(pick))
    shp.append((int(rng.integers(300, 1300)), int(rng.integers(300, 1300))))
lab_d = torch.from_numpy(np.concatenate(labs).astype(np.float32)).to(dev)
offs_d = torch.tensor(offs, dtype=torch.int32, device=dev)
shp_d = torch.tensor(shp, dtype=torch.int32, device=dev)
iouv = torch.linspace(0.5, 0.95, 10, device=dev)
ms = timeit(lambda: ops.eval_match(nout.dets, nout.counts, lab_d, offs_d, 25, shp_d, (640, 640), iouv))
report("eval_match (300 dets x 25 labels)", ms, B * (300 * 24 * 2 + 300 * 10 + 25 * 44), "latency-bound: one block per image")
srcs = [torch.from_numpy(synth.image_u8(h, w, 50 + i)).to(dev) for i, (h, w) in enumerate(sizes)]
plan = ops.LetterboxPlan(srcs, (640, 640), mode="dataset")
out = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
ms = timeit(lambda: plan.run(out))
report("dataset ingest mixed 480-1280 f32", ms, src_bytes + B * 3 * 640 * 640 * 4, "load_resized_image + PadIfNeeded + /255")
print(json.dumps({"batch": B, "peak_GBps": PEAK, "kernels": res}))
