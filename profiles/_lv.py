import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops, synth
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, g, B in (("P3", 80, 64), ("P4", 40, 256), ("P5", 20, 1024)):
    cfg = ops.head_cfg("v5", 80, [synth.V5_ANCHORS[0]], [8.0], [(g, g)])
    lv = [torch.randn(B, 255, g, g, device=dev)]
    ms = timeit(lambda: ops.detect_decode(cfg, lv))
    nb = 2 * lv[0].numel() * 4
    print(f"{name} B={B}: {ms*1e3:8.1f} us  {nb/ms/1e6:8.1f} GB/s")
import time
for B in (1, 64):
    cfg = ops.head_cfg("v5", 80, [synth.V5_ANCHORS[0]], [8.0], [(80, 80)])
    lv = [torch.randn(B, 255, 80, 80, device=dev)]
    ms = timeit(lambda: ops.detect_decode(cfg, lv), iters=200)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): ops.detect_decode(cfg, lv)
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"B={B}: event {ms*1e3:.1f} us/call, host-side issue time {(t1-t0)/200*1e6:.1f} us/call")
lvz = [torch.zeros(64, 255, 80, 80, device=dev)]
print("zeros input:", timeit(lambda: ops.detect_decode(cfg, lvz)) * 1e3, "us")
