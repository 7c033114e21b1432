#!/usr/bin/env python
"""BASELINE config 1 (demo path) latency: one 1080x810 BGR frame -> ImageProcessor.preprocess, a
(1,25200,85) prediction -> ImageProcessor.postprocess (nms + un-letterbox of image 0), through the
drop-in classes, next to the CPU reference port.  Wall clock with a device synchronisation on
both sides, median of 200 frames (the reference's own demo prints unsynchronised times)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import ref_port
from vision_kit_b200 import ops
from tests import synth
from vision_kit_b200.processing import ImageProcessor

dev = torch.device("cuda:0")
frame = synth.image_u8(1080, 810, 1)          # stands in for assets/bus.jpg (1080x810)
lv = synth.head_logits(1, seed=2, clusters=20)
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
pred = ops.detect_decode(cfg, [torch.from_numpy(x).to(dev) for x in lv])
pred_cpu = pred.cpu()
ip = ImageProcessor(auto=False)


def med(fn, n=200, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


res = {}
res["gpu_preprocess_ms"] = med(lambda: ip.preprocess(frame))
res["gpu_postprocess_ms"] = med(lambda: ip.postprocess(pred))
x, _ = ip.preprocess(frame)
dets = ip.postprocess(pred)
res["detections"] = int(dets.shape[0])
res["cpu_preprocess_ms"] = med(lambda: ref_port.preprocess(frame, (640, 640)), n=50, warm=3)


def cpu_post():
    d = ref_port.nms(pred_cpu)[0]
    g = ops.letterbox_geometry(1080, 810, (640, 640))
    d[:, [0, 2]] -= g.pad_w
    d[:, [1, 3]] -= g.pad_h
    d[:, :4] /= g.ratio
    return d


res["cpu_postprocess_ms"] = med(cpu_post, n=50, warm=3)
res["cpu_threads"] = torch.get_num_threads()
exp, _ = ref_port.preprocess(frame, (640, 640))
res["preprocess_bit_exact"] = bool(torch.equal(x.cpu(), exp))
res["postprocess_bit_exact"] = bool(torch.equal(dets.cpu(), cpu_post()))
print(json.dumps(res))
