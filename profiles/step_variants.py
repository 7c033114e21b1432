import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tests import synth
from vision_kit_b200.pipeline import DetectPipeline
dev = torch.device("cuda:0")
B = 64
ident = list(torch.from_numpy(synth.images_u8(B, 640, 640, seed=0)).to(dev))
lv2 = [torch.from_numpy(x).to(dev) for x in synth.head_logits(B, seed=2, clusters=20)]
for overlap, fork in ((True, False), (True, True), (False, False)):
    pipe = DetectPipeline("v5", batch=B, device=dev, overlap=overlap, fork_preprocess=fork)
    pipe.plan_sources(ident); pipe.capture(lv2)
    for _ in range(10): pipe.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): pipe.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"VK_NMS_T={os.environ.get('VK_NMS_T','-')} overlap={overlap} fork={fork}: {e0.elapsed_time(e1)/200*1e3:.1f} us/step")
