#!/usr/bin/env python
"""How much of a config-2 step is the gap between two graph launches: the captured step (one graph per batch) against
graphs that hold 2 / 4 consecutive steps.  python profiles/graph_depth.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from vision_kit_b200 import _lib
from vision_kit_b200.pipeline import DetectPipeline
dev = torch.device("cuda:0")
B = 64
ident = list(torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev))
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
pipe = DetectPipeline("v5", batch=B, device=dev, overlap=True)
pipe.plan_sources(ident); pipe.capture(lv)
L = pipe._lib
sp = lambda s: C.c_void_p(s.cuda_stream)


def body(s, main):
    late = pipe.nms_fork == "after_preprocess"
    if late:
        _lib.check("lb", L.vk_letterbox_batch(*pipe._lb_args, sp(main)))
    pipe.side.wait_stream(main)
    _lib.check("nms", L.vk_nms_batched(*pipe._nms_args[s ^ 1], sp(pipe.side)))
    if not late:
        _lib.check("lb", L.vk_letterbox_batch(*pipe._lb_args, sp(main)))
    _lib.check("filter", L.vk_decode_filter(pipe._cfg_ref, C.cast(pipe._lv_arr, C.c_void_p), pipe._lv_dt, B, pipe._conf,
                                            pipe._ml, pipe._mask_p, pipe._kernel, C.byref(pipe._cs[s]), sp(main)))
    main.wait_stream(pipe.side)


def timeit(fn, per, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * per) * 1e3


print(f"1 step per graph : {timeit(pipe.replay, 1):.1f} us/step")
for depth in (2, 4, 8):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for k in range(depth):
            body(k & 1, main)
    print(f"{depth} steps per graph: {timeit(g.replay, depth, 200 // depth):.1f} us/step")
