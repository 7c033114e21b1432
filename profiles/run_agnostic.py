#!/usr/bin/env python
"""Launches the eval-mode agnostic NMS a few times (for ncu captures): python profiles/run_agnostic.py [batch]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import _lib, ops
from tests import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v7", 80, synth.V7_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x[:B]).to(dev) for x in synth.head_logits(64, seed=4, clusters=20)]
rows, segs = ops.head_rows(cfg), _lib.lib().vk_decode_filter_segments(C.byref(cfg))
buf = ops.CandBuf.alloc(B, rows, segs, 80, ops.default_cap(segs, 80, True), dev, top_list=True, list_cap=2 * ops.LIST_CAP)
ops.decode_filter(cfg, lv, 0.001, True, buf=buf)
out = ops.nms_batched(buf, 0.6, True)
for _ in range(3):
    ops.nms_batched(buf, 0.6, True, out=out)
torch.cuda.synchronize()
print("ok", int(out.counts.sum()))
