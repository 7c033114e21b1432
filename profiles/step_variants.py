#!/usr/bin/env python
"""Captured step of config 3 on one GPU with both fork points of the NMS branch, VK_BATCH images per step (variant
libraries via VK_B200_LIB): python profiles/step_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from vision_kit_b200.pipeline import DetectPipeline
dev = torch.device("cuda:0")
B = int(os.environ.get("VK_BATCH", "64"))
ident = list(torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev))
lv2 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
EVAL = dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)
for fork in ("start", "after_preprocess"):
    pipe = DetectPipeline("v5", batch=B, device=dev, overlap=True, nms_fork=fork, **EVAL)
    pipe.plan_sources(ident); pipe.capture(lv2)
    for _ in range(10): pipe.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): pipe.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"config 3 B={B} nms_fork={fork:16s}: {e0.elapsed_time(e1)/200*1e3:.1f} us/step", flush=True)
    del pipe
