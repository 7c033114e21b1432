#!/usr/bin/env python
"""Eval-mode NMS launches at 256 images (ncu launch list target): python profiles/run_nms256.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
buf = ops.decode_filter(cfg, lv, 0.001, True)
out = ops.nms_batched(buf, 0.6)
for _ in range(3):
    ops.nms_batched(buf, 0.6, out=out)
torch.cuda.synchronize()
print("ok", int(out.counts.sum()))
