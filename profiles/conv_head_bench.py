#!/usr/bin/env python
"""Fused Detect head (vk_conv_decode_filter: tcgen05 1x1 conv + decode + filter) against the unfused
path (torch/cuDNN 1x1 conv in fp32 and with TF32 allowed -> vk_decode_filter), YOLOv5s head widths,
B = 64 at 640: python profiles/conv_head_bench.py > gpurun_out/conv_head.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
PERSISTENT = os.environ.get("VK_CONV_MODE", "1") == "1"   # 1 = persistent warp-specialised kernel (default), 0 = tile kernel
dev = torch.device("cuda:0")
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
res = {}


def timeit(fn, iters=10, warm=3):
    """Device time per call: the calls are captured into one CUDA graph (the eager loop is host-bound)."""
    for _ in range(warm):
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


for name, cins in (("yolov5s", (128, 256, 512)), ("yolov5x", (320, 640, 1280))):
    cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
    g = torch.Generator().manual_seed(1)
    feats = [torch.randn(B, c, n, n, generator=g).to(dev) for c, (n, _) in zip(cins, grids)]
    ws = [(torch.randn(255, c, 1, 1, generator=g) * (1.2 / c ** 0.5)).to(dev) for c in cins]
    bs = []
    for _ in cins:
        b = torch.randn(255, generator=g) * 0.5
        b.view(3, 85)[:, 4] -= 3.0
        b.view(3, 85)[:, 5:] -= 1.5
        bs.append(b.to(dev))
    in_bytes = sum(f.numel() * 4 for f in feats)
    flops = 2 * 255 * sum(f.numel() for f in feats)
    for mode, conf, ml in (("demo", 0.25, False), ("eval", 0.001, True)):
        buf = ops.conv_decode_filter(cfg, feats, ws, bs, conf, ml, persistent=PERSISTENT)
        t_fused = timeit(lambda: ops.conv_decode_filter(cfg, feats, ws, bs, conf, ml, buf=buf, persistent=PERSISTENT))
        row = {"fused_us": round(t_fused, 1), "candidates_per_img": int(buf.counts.sum()) // B,
               "in_MB": round(in_bytes / 1e6, 1), "GFLOP": round(flops / 1e9, 2),
               "fused_TFLOPs": round(flops / t_fused / 1e6, 1), "fused_in_GBps": round(in_bytes / t_fused / 1e3, 1)}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            conv = lambda: [torch.nn.functional.conv2d(f, w, b) for f, w, b in zip(feats, ws, bs)]
            lv = conv()
            buf2 = ops.decode_filter(cfg, lv, conf, ml)
            row["cudnn_%s_candidates_per_img" % ("tf32" if tf32 else "fp32")] = int(buf2.counts.sum()) // B
            t_conv = timeit(conv)
            t_both = timeit(lambda: ops.decode_filter(cfg, conv(), conf, ml, buf=buf2))
            row["cudnn_%s_conv_us" % ("tf32" if tf32 else "fp32")] = round(t_conv, 1)
            row["cudnn_%s_conv+filter_us" % ("tf32" if tf32 else "fp32")] = round(t_both, 1)
        res[f"{name} {mode}"] = row
        print(f"{name} {mode}: {row}", file=sys.stderr)
print(json.dumps({"batch": B, "results": res}))
