#!/usr/bin/env python
"""Phase breakdown of nms_kernel from clock64 stamps (profiling build: python __graft_entry__.py --profiling):
    VK_B200_LIB=vision_kit_b200/libvk_b200_prof.so python profiles/nms_phases.py
stamp 0 start, 1 set-up done, then per stage (2k, 2k+1) = (selection + sort + box fetch done, chunks done), 30 end."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vision_kit_b200 import _lib, ops
from tests import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
L = _lib.lib()
L.vkdbg_nms_timing.argtypes = [C.c_void_p]
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
lv = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
res = {}
for mode, conf, ml, iou, agn in (("demo", 0.25, False, 0.45, False), ("eval", 0.001, True, 0.6, False),
                                 ("eval agnostic", 0.001, True, 0.6, True)):
    buf = ops.decode_filter(cfg, lv, conf, ml)
    out = ops.nms_batched(buf, iou, agn)
    stamps = torch.zeros((B, 32), dtype=torch.int64, device=dev)
    L.vkdbg_nms_timing(C.c_void_p(stamps.data_ptr()))
    ops.nms_batched(buf, iou, agn, out=out)
    torch.cuda.synchronize()
    L.vkdbg_nms_timing(None)
    s = stamps.cpu().numpy()
    clk = 1.9     # GHz, nominal under light load
    names = ["setup", "radix", "compact", "sort", "fetch", "c0 group", "c0 kept", "c0 pred", "c0 resolve", "c0 output", "chunks"]
    names += ["s1 " + x for x in names[1:]]
    rows = []
    for b in range(B):
        t = s[b]
        marks, prev = {}, int(t[0])
        for k, nm in enumerate(names, start=1):
            if t[k] > 0:
                marks[nm] = round((int(t[k]) - prev) / clk / 1e3, 2)
                prev = int(t[k])
        rows.append({"n": int(t[31] >> 32), "processed": int(t[31] & 0xffffffff), "total_us": round((int(t[30]) - int(t[0])) / clk / 1e3, 1),
                     "phase_us": marks})
    tot = np.array([r["total_us"] for r in rows])
    order = np.argsort(tot)
    print(f"{mode:14s} per-image CTA time: mean {tot.mean():.1f} us, max {tot.max():.1f} us, min {tot.min():.1f}", file=sys.stderr)
    print(f"   median image: {rows[int(order[B // 2])]}", file=sys.stderr)
    print(f"   slowest image: {rows[int(order[-1])]}", file=sys.stderr)
    res[mode] = {"mean_us": float(tot.mean()), "max_us": float(tot.max()), "median": rows[int(order[B // 2])], "slowest": rows[int(order[-1])]}
print(json.dumps(res))
