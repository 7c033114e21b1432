#!/usr/bin/env python
"""Launches the fused conv head a few times (for ncu captures): python profiles/run_conv.py [batch] [conf] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import _lib, ops
from tests import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
conf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
PERSISTENT = (int(sys.argv[3]) if len(sys.argv) > 3 else 1) == 1
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
g = torch.Generator().manual_seed(1)
cins = (128, 256, 512)
feats = [torch.randn(B, c, n, n, generator=g).to(dev) for c, (n, _) in zip(cins, grids)]
ws = [(torch.randn(255, c, 1, 1, generator=g) * (1.2 / c ** 0.5)).to(dev) for c in cins]
bs = []
for _ in cins:
    b = torch.randn(255, generator=g) * 0.5
    b.view(3, 85)[:, 4] -= 3.0
    b.view(3, 85)[:, 5:] -= 1.5
    bs.append(b.to(dev))
for _ in range(3):
    buf = ops.conv_decode_filter(cfg, feats, ws, bs, conf, conf < 0.05, persistent=PERSISTENT)
torch.cuda.synchronize()
print("ok", int(buf.counts.sum()))
