#!/usr/bin/env python
"""BASELINE config 5 in miniature, on real GPUs over NCCL (run under torchrun, one rank per GPU):
mixed-resolution sources -> letterbox -> fused decode+filter -> NMS on each rank's shard, then the
mAP-eval all-gather of padded detections; every rank checks the gathered result against the same
pipeline run locally on ALL images.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_eval_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from vision_kit_b200 import dist as vkd, ops
from tests import synth

N_IMAGES = 22          # not a multiple of the world size: shards differ by one image


def run(images, logits, dev):
    srcs = [torch.from_numpy(i).to(dev) for i in images]
    x, rps = ops.letterbox_batch(srcs, (640, 640), swap_rb=True)
    lv = [torch.from_numpy(l).to(dev) for l in logits]
    grids = [(640 // s, 640 // s) for s in synth.STRIDES]
    cfg = ops.head_cfg("v7", 80, synth.V7_ANCHORS, synth.STRIDES, grids)
    out = ops.nms_batched(ops.decode_filter(cfg, lv, 0.001, True), 0.6)
    return x, out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    sizes = synth.mixed_sizes(N_IMAGES, seed=3)
    images = [synth.image_u8(h, w, 200 + i) for i, (h, w) in enumerate(sizes)]
    logits = synth.head_logits(N_IMAGES, seed=5, clusters=15)
    lo, hi = vkd.shard_range(N_IMAGES, rank, world)
    x, out = run(images[lo:hi], [l[lo:hi] for l in logits], dev)
    dets, counts = vkd.allgather_detections(out.dets, out.counts, N_IMAGES)
    x_all, ref = run(images, logits, dev)
    ok = torch.equal(dets, ref.dets) and torch.equal(counts, ref.counts) and torch.equal(x, x_all[lo:hi])
    ok = ok and int(counts.sum()) > 0
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_eval_check world={world} images={N_IMAGES} detections={int(counts.sum())} "
              f"{'OK' if int(flag) else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
