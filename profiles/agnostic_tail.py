#!/usr/bin/env python
"""Per-image CTA time of the eval-mode agnostic NMS on bench.py's config-4 data of every rank (profiling build:
python -c "import __graft_entry__ as g; g.build_variant('prof', ['-DVK_NMS_PROFILE'])"):
    VK_B200_LIB=vision_kit_b200/libvk_b200_prof.so python profiles/agnostic_tail.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vision_kit_b200 import _lib, ops
from tests import synth

B = 32
dev = torch.device("cuda:0")
L = _lib.lib()
L.vkdbg_nms_timing.argtypes = [C.c_void_p]
grids = [(640 // s, 640 // s) for s in synth.STRIDES]
cfg = ops.head_cfg("v7", 80, synth.V7_ANCHORS, synth.STRIDES, grids)
for rank in range(8):
    lv = [torch.from_numpy(x[:B]).to(dev) for x in synth.head_logits(64, seed=4 + 16 * rank, clusters=20)]
    for agn in (False, True):
        rows, segs = ops.head_rows(cfg), L.vk_decode_filter_segments(C.byref(cfg))
        buf = ops.CandBuf.alloc(B, rows, segs, 80, ops.default_cap(segs, 80, True), dev, top_list=True,
                                list_cap=ops.LIST_CAP * (2 if agn else 1))      # (DetectPipeline doubles the list for agnostic NMS)
        ops.decode_filter(cfg, lv, 0.001, True, buf=buf)
        out = ops.nms_batched(buf, 0.6, agn)
        stamps = torch.zeros((B, 32), dtype=torch.int64, device=dev)
        L.vkdbg_nms_timing(C.c_void_p(stamps.data_ptr()))
        ops.nms_batched(buf, 0.6, agn, out=out)
        torch.cuda.synchronize()
        L.vkdbg_nms_timing(None)
        s = stamps.cpu().numpy()
        tot = (s[:, 30] - s[:, 0]) / 1.9e3
        n, proc = (s[:, 31] >> 32), (s[:, 31] & 0xffffffff)
        w = int(np.argmax(tot))
        print(f"rank {rank} {'agnostic' if agn else 'class   '}: per-image us mean {tot.mean():6.1f} max {tot.max():6.1f} (image {w}: n={int(n[w])} processed={int(proc[w])}, "
              f"stamps {[round((int(s[w, k]) - int(s[w, 0])) / 1.9e3, 1) for k in range(1, 12) if s[w, k] > 0]})", flush=True)
