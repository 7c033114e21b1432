#!/usr/bin/env python
"""Phase breakdown of nms_image_kernel from in-kernel clock64 stamps (GPU box only):
    python profiles/nms_phases.py            # bench workload (B=64, demo NMS)"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import _lib, ops, synth
from vision_kit_b200.pipeline import DetectPipeline
B = 64
dev = torch.device("cuda:0")
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
mode = sys.argv[1] if len(sys.argv) > 1 else "demo"
kw = dict(conf_thres=0.25, iou_thres=0.45) if mode == "demo" else dict(conf_thres=0.001, iou_thres=0.6, multi_label=True)
pipe = DetectPipeline("v5", batch=B, device=dev, **kw)
buf = torch.zeros((B, 32), dtype=torch.int64, device=dev)
lib = _lib.lib()
for _ in range(3):
    pipe.postprocess(lv)
torch.cuda.synchronize()
lib.vkdbg_nms_timing.argtypes = [C.c_void_p]
assert lib.vkdbg_nms_timing(C.c_void_p(buf.data_ptr())) == 0
pipe.postprocess(lv)
torch.cuda.synchronize()
lib.vkdbg_nms_timing(C.c_void_p(0))
t = buf.cpu().numpy()
names = ["load counts+A1", "A2 select", "A3 compaction", "A4 sort", "A5 init+chunk0 load", "chunk0 phases 1-2",
         "chunk0 resolve", "chunk0 output", "remaining chunks", "tail"]
d = (t[:, 1:11] - t[:, 0:10])
print(f"mode={mode}  n: mean {t[:,11].mean():.0f} max {t[:,11].max()}  kept: mean {t[:,12].mean():.0f}")
print(f"total cycles per CTA: mean {(t[:,10]-t[:,0]).mean():.0f} max {(t[:,10]-t[:,0]).max()}")
for i, nme in enumerate(names):
    print(f"  {nme:24s} mean {d[:, i].mean():9.0f}  max {d[:, i].max():9.0f} cycles")
print(f"  chunk0 phase 1 only: mean {(t[:,13]-t[:,5]).mean():.0f}; resolve rounds chunk0: mean {t[:,14].mean():.1f} max {t[:,14].max()}")
print(f"  resolve: load p {(t[:,15]-t[:,6]).mean():.0f}, rounds {(t[:,16]-t[:,15]).mean():.0f}, exit->sync {(t[:,7]-t[:,16]).mean():.0f}")
print("kernel span (first start -> last end):", t[:, 10].max() - t[:, 0].min(), "cycles")
