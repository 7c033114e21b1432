#!/usr/bin/env python
"""Device time of the demo-mode filter and NMS kernels of whichever library VK_B200_LIB names (tuning builds):
    VK_B200_LIB=vision_kit_b200/libvk_b200_<variant>.so python profiles/rows_bench.py [tag]"""
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


def timeit(fn, iters=20):
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


buf = ops.decode_filter(cfg, lv, 0.25, False)
t_f = timeit(lambda: ops.decode_filter(cfg, lv, 0.25, False, buf=buf))
out = ops.nms_batched(buf, 0.45)
t_n = timeit(lambda: ops.nms_batched(buf, 0.45, out=out))
pred = ops.detect_decode(cfg, lv)
bufp = ops.filter_pred(pred, 0.25, False)
t_p = timeit(lambda: ops.filter_pred(pred, 0.25, False, buf=bufp))
print(f"{tag:28s} decode_filter demo {t_f:7.1f} us   filter_pred demo {t_p:7.1f} us   nms demo {t_n:7.1f} us   ({int(buf.counts.sum())} cand, {int(out.counts.sum())} dets)")
