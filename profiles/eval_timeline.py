#!/usr/bin/env python
"""Where the kernels of an overlapped eval step (config 3) sit in time: the captured step issued call by call on the
two streams with CUDA events around each call, offsets from the start of the step averaged over the steps.
python profiles/eval_timeline.py [nms_fork] [side_priority]"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from vision_kit_b200 import _lib
from vision_kit_b200.pipeline import DetectPipeline
dev = torch.device("cuda:0")
B = 64
fork = sys.argv[1] if len(sys.argv) > 1 else "start"
prio = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ident = list(torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev))
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
pipe = DetectPipeline("v5", batch=B, device=dev, overlap=True, nms_fork=fork, side_priority=prio,
                      conf_thres=0.001, iou_thres=0.6, multi_label=True)
pipe.plan_sources(ident); pipe.capture(lv)
L = pipe._lib
main, side = torch.cuda.current_stream(), pipe.side
sp = lambda s: C.c_void_p(s.cuda_stream)
N = 30
acc = [0.0] * 6
for it in range(N + 5):
    s = it & 1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    torch.cuda.synchronize()
    ev[0].record(main)
    late = fork == "after_preprocess"

    def nms():
        side.wait_stream(main)
        ev[1].record(side)
        _lib.check("nms", L.vk_nms_batched(*pipe._nms_args[s ^ 1], sp(side)))
        ev[2].record(side)
    if not late:
        nms()
    ev[3].record(main)
    _lib.check("lb", L.vk_letterbox_batch(*pipe._lb_args, sp(main)))
    ev[4].record(main)
    if late:
        nms()
    _lib.check("filter", L.vk_decode_filter(pipe._cfg_ref, C.cast(pipe._lv_arr, C.c_void_p), pipe._lv_dt, B, pipe._conf,
                                            pipe._ml, pipe._mask_p, pipe._kernel, C.byref(pipe._cs[s]), sp(main)))
    ev[5].record(main)
    main.wait_stream(side)
    ev[6].record(main)
    torch.cuda.synchronize()
    if it >= 5:
        for k in range(6):
            acc[k] += ev[0].elapsed_time(ev[k + 1]) * 1e3 / N
names = ["nms start", "nms end", "letterbox start", "letterbox end", "filter end", "step end"]
print(f"nms_fork={fork} side_priority={prio} (us from the start of the step, eager issue): " +
      ", ".join(f"{n} {v:.0f}" for n, v in zip(names, acc)))
