"""__graft_entry__.smoke(): one small pass of the hot path on cuda:0 checked against the
oracle (the only place outside tests/ and bench.py that imports oracle/, as the checker)."""
from __future__ import annotations

import numpy as np
import torch


def run() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device")
    from oracle import ref_port
    from tests import synth
    from . import _lib, ops
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    n0 = _lib.launch_count()
    # letterbox: two sources of different sizes -> one (2,3,640,640) batch, bit-exact vs cv2 path
    imgs = [synth.image_u8(720, 1280, 1), synth.image_u8(500, 375, 2)]
    out, rps = ops.letterbox_batch([torch.from_numpy(i).to(dev) for i in imgs], (640, 640), swap_rb=True)
    for i, im in enumerate(imgs):
        exp, (ratio, pad) = ref_port.preprocess(im, (640, 640))
        assert torch.equal(out[i].cpu(), exp[0]), "letterbox mismatch"
        assert rps[i][0] == ratio
    # decode + filter + NMS, fused, vs the oracle on the same logits
    lv = synth.head_logits(2, seed=7, clusters=15)
    grids = [(640 // s, 640 // s) for s in synth.STRIDES]
    cfg = ops.head_cfg("v5", 80, synth.V5_ANCHORS, synth.STRIDES, grids)
    dev_lv = [torch.from_numpy(x).to(dev) for x in lv]
    res = ops.nms_batched(ops.decode_filter(cfg, dev_lv, 0.25, False), 0.45, want_keep=True)
    pred = ops.detect_decode(cfg, dev_lv)
    outs, keeps = ref_port.nms(pred.cpu(), return_keep=True)
    for i in range(2):
        k = int(res.counts[i])
        assert k == outs[i].shape[0] and k > 0, "detection count mismatch"
        assert np.array_equal(res.keep[i, :k].cpu().numpy(), keeps[i].numpy()), "keep index mismatch"
        assert np.array_equal(res.dets[i, :k].cpu().numpy(), outs[i].numpy()), "detections mismatch"
    exp_pred, _ = ref_port.detect_decode([torch.from_numpy(x) for x in lv], synth.V5_ANCHORS, synth.STRIDES, "v5")
    np.testing.assert_allclose(pred.cpu().numpy(), exp_pred.numpy(), rtol=1e-5, atol=1e-6)
    torch.cuda.synchronize()
    print(f"smoke ok: {_lib.launch_count() - n0} kernel launches of libvk_b200.so, "
          f"{int(res.counts.sum())} detections, parity with the oracle")
