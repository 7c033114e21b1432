#!/usr/bin/env python
"""Device time of the eval-mode (dense) fused filter of whichever library VK_B200_LIB names (tuning builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("VK_B200_LIB", "default"))
B = 64
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
lvh = [t.half() for t in lv]


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
for name, levels, conf, ml, k in (("eval ML f32", lv, 0.001, True, "dense"), ("eval ML f16", lvh, 0.001, True, "dense"),
                                  ("eval ML f32 one-pass", lv, 0.001, True, "dense_onepass"), ("best-class dense f32", lv, 0.01, False, "dense")):
    buf = ops.decode_filter(cfg, levels, conf, ml, kernel=k)
    out.append(f"{name} {timeit(lambda: ops.decode_filter(cfg, levels, conf, ml, buf=buf, kernel=k)):7.1f} us")
# an odd grid (608: 76 / 38 / 19): the pairs kernel loads the two rows of a lane separately
grids6 = [(608 // s, 608 // s) for s in synth.STRIDES]
cfg6 = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids6)
lv6 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20, img=608)]
for name, k in (("608 eval ML f32", "dense"), ("608 one-pass", "dense_onepass")):
    buf = ops.decode_filter(cfg6, lv6, 0.001, True, kernel=k)
    out.append(f"{name} {timeit(lambda: ops.decode_filter(cfg6, lv6, 0.001, True, buf=buf, kernel=k)):7.1f} us")
print(f"{tag:24s} " + "   ".join(out) + f"   ({int(buf.counts.sum())} cand)")
