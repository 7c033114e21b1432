#!/usr/bin/env python
"""Device time of the eval-mode NMS (select pass + per-image kernel) of whichever library VK_B200_LIB names:
python profiles/nms_bench.py [tag]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("VK_B200_LIB", "default"))
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


out = []
for B in (32, 64, 256):
    lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
    buf = ops.decode_filter(cfg, lv, 0.001, True)
    for agn in (False, True):
        o = ops.nms_batched(buf, 0.6, agnostic=agn)
        out.append(f"B={B}{' agnostic' if agn else ''} {timeit(lambda: ops.nms_batched(buf, 0.6, agnostic=agn, out=o)):7.1f} us")
    del lv, buf
print(f"{tag:16s} nms eval: " + "   ".join(out))
