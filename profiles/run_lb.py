#!/usr/bin/env python
"""Launches the general letterbox kernel a few times (for ncu captures):
    python profiles/run_lb.py [mixed|up|down] [f32|bf16] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_kit_b200 import ops
from tests import synth
kind = sys.argv[1] if len(sys.argv) > 1 else "mixed"
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda:0")
sizes = {"mixed": synth.mixed_sizes(B, seed=5), "up": [(480, 480)] * B, "down": [(1280, 1280)] * B}[kind]
srcs = [torch.from_numpy(synth.image_u8(h, w, 50 + i)).to(dev) for i, (h, w) in enumerate(sizes)]
plan = ops.LetterboxPlan(srcs, (640, 640))
out = torch.empty((B, 3, 640, 640), dtype=dt, device=dev)
for _ in range(3):
    plan.run(out, swap_rb=True)
torch.cuda.synchronize()
print("ok", float(out.float().mean()))
