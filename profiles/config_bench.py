#!/usr/bin/env python
"""Device-resident pipeline throughput of the other BASELINE configs on ONE GPU's shard (CUDA events,
one CUDA-graph launch per step, NMS beside the next batch on a side stream unless stated): python profiles/config_bench.py > gpurun_out/configs.json
  config 3: YOLOv5x-shaped head, eval NMS (conf 0.001, iou 0.6, multi-label), identity letterbox
  config 4: YOLOv7 decode order, eval NMS, agnostic vs class-aware
  config 5: mixed 480-1280 sources -> 640 letterbox, YOLOv7 decode, eval NMS
(head tensors have the same shape for every model size: only the conv in-channels differ)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from vision_kit_b200.pipeline import DetectPipeline
dev = torch.device("cuda:0")
res = {}


def run(name, variant, B, srcs, lv, steps=40, warm=5, overlap=True, fork=False, **kw):
    pipe = DetectPipeline(variant, batch=B, device=dev, overlap=overlap, fork_preprocess=fork, **kw)
    pipe.plan_sources(srcs)
    pipe.capture(lv)

    def step():
        pipe.replay()
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pipe.flush() if overlap else None
    res[name] = {"batch": B, "ms_per_step": round(ms, 4), "images_per_s": round(B / ms * 1e3),
                 "detections_per_image": int(pipe.out.counts.sum()) // B}
    print(f"{name:46s} B={B:3d} {ms*1e3:8.1f} us/step {B/ms*1e3:10.0f} img/s", file=sys.stderr)


B = 64
ident = list(torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev))
lv3 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=3, clusters=20)]
ev = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)
run("config 3 (v5, eval NMS, identity letterbox)", "v5", B, ident, lv3, **ev)
lv4 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=4, clusters=20)]
run("config 4 (v7, eval NMS, class-aware)", "v7", B, ident, lv4, **ev)
run("config 4 (v7, eval NMS, agnostic)", "v7", B, ident, lv4, agnostic=True, **ev)
sizes = synth.mixed_sizes(B, seed=5)
mixed = [torch.from_numpy(synth.image_u8(h, w, 50 + i)).to(dev) for i, (h, w) in enumerate(sizes)]
run("config 5 (mixed 480-1280 letterbox, v7, eval NMS)", "v7", B, mixed, lv4, **ev)
lv2 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
run("config 2 (v5, demo NMS, identity letterbox)", "v5", B, ident, lv2, conf_thres=0.25, iou_thres=0.45)
run("config 2, letterbox forked beside the filter", "v5", B, ident, lv2, fork=True, conf_thres=0.25, iou_thres=0.45)
run("config 2, single stream", "v5", B, ident, lv2, overlap=False, conf_thres=0.25, iou_thres=0.45)
run("config 3, letterbox forked beside the filter", "v5", B, ident, lv3, fork=True, **ev)
run("config 3, single stream", "v5", B, ident, lv3, overlap=False, **ev)
print(json.dumps(res))
