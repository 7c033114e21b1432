#!/usr/bin/env python
"""Launches each secondary kernel a few times (for ncu captures):
    python profiles/run_one.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
sizes = synth.mixed_sizes(B, seed=5)
srcs = [torch.from_numpy(synth.image_u8(h, w, 50 + i)).to(dev) for i, (h, w) in enumerate(sizes)]
plan = ops.LetterboxPlan(srcs, (640, 640))
out = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
for _ in range(3):
    plan.run(out, swap_rb=True)
    pred = ops.detect_decode(cfg, lv)
    buf = ops.decode_filter(cfg, lv, 0.001, True)
    buf2 = ops.filter_pred(pred, 0.001, True)
    res = ops.nms_batched(buf, 0.6)
torch.cuda.synchronize()
print("ok", int(res.counts.sum()))
