"""Drop-in for ``vision_kit.utils.image_proc`` (reference utils/image_proc.py): the same
``resize`` / ``scale_coords`` / ``nms`` names, arguments, defaults, return structure and
in-place behaviour, executed by the sm_100a kernels of libvk_b200.so.

Differences a caller can observe (deliberate):
  * ``nms`` has no wall-clock time limit (utils/image_proc.py:109,183-185 silently drops
    images when it fires) and its ``n > max_nms`` pre-cut breaks score ties by candidate
    order (stable), where torch's CPU argsort leaves the order unspecified.
  * tensors must live on a CUDA device; there is no CPU path.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import _lib, ops

MAX_NMS = 30000     # utils/image_proc.py:108


def resize(img_sz: Tuple[int, int], img: np.ndarray, stride: int = 32, letterbox: bool = True,
           scaleup: bool = True, auto: bool = False,
           color: Tuple[int, int, int] = (114, 114, 114)):
    """utils/image_proc.py:12-60.  numpy HWC uint8 in, numpy HWC uint8 out (the GPU does the
    resize + border; the copies are the price of the numpy signature -- batched device
    callers use ``ops.letterbox_batch``).  Returns ``(img, (ratio, (dw, dh)))``."""
    src = torch.from_numpy(np.ascontiguousarray(img)).cuda(non_blocking=False)
    out, rps = ops.letterbox_batch([src], img_sz, stride, letterbox, scaleup, auto, color,
                                   swap_rb=False, dtype=torch.uint8)
    ratio, pad = rps[0]
    if not letterbox:
        pad = (int(pad[0]), int(pad[1]))          # :46-47,60 -- ints, not halved
    return out[0].cpu().numpy(), (ratio, pad)


def scale_coords(img1_shape, coords: torch.Tensor, img0_shape, ratio_pad=None) -> torch.Tensor:
    """utils/image_proc.py:63-80: rescales ``coords`` IN PLACE (the only caller,
    core/eval/det_evaluator.py:154-168, relies on that) and returns the clone taken before
    the edit, exactly like the reference."""
    converted = coords.clone()
    if ratio_pad is None:                                             # :67-71
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    ops.scale_coords_(coords, pad[0], pad[1], gain, True, clip_hw=img0_shape)   # :76-79
    return converted


def _append_labels(prediction: torch.Tensor, labels, conf_thres: float) -> torch.Tensor:
    """A-priori labels (utils/image_proc.py:122-128): rows [box, obj=1, one-hot cls] appended
    after the image's own rows.  Images with fewer labels get rows with obj=0, which the
    ``obj > conf`` filter drops."""
    bs, rows, no = prediction.shape
    lmax = max((len(l) for l in labels), default=0)
    if lmax == 0:
        return prediction
    extra = torch.zeros((bs, lmax, no), dtype=prediction.dtype, device=prediction.device)
    for i, lb in enumerate(labels):
        if len(lb):
            lb = torch.as_tensor(lb, dtype=prediction.dtype, device=prediction.device)
            k = lb.shape[0]
            extra[i, :k, :4] = lb[:, 1:5]
            extra[i, :k, 4] = 1.0
            extra[i, torch.arange(k, device=prediction.device), lb[:, 0].long() + 5] = 1.0
    return torch.cat((prediction, extra), 1).contiguous()


_BUF_CACHE = {}


def _cand_buffer(dev, stream, bs: int, rows: int, nc: int, multi_label: bool, top_list: bool) -> ops.CandBuf:
    """The worst-case candidate buffer of a drop-in call is large (segs * 64 * nc slots per image in
    multi-label mode), so one per (device, stream, shape) is kept and reused; work on one stream is
    ordered, which makes the reuse safe."""
    key = (dev, stream, bs, rows, nc, multi_label, top_list)
    buf = _BUF_CACHE.get(key)
    if buf is None:
        if len(_BUF_CACHE) >= 4:
            _BUF_CACHE.clear()
        segs = _lib.lib().vk_filter_segments(rows)
        buf = ops.CandBuf.alloc(bs, rows, segs, nc, ops.default_cap(segs, nc, multi_label), dev, top_list=top_list)
        _BUF_CACHE[key] = buf
    return buf


def _run_nms(prediction: torch.Tensor, conf_thres: float, iou_thres: float, classes, agnostic: bool,
             multi_label: bool, labels, max_det: int, max_nms: int, want_keep: bool = False):
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    _lib.require_cuda(prediction, "nms(prediction)")
    pred = prediction
    if pred.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        pred = pred.float()
    if labels:
        if len(labels) != pred.shape[0]:           # the reference indexes labels[xi] (:122-123) and fails
            raise IndexError(f"nms: {len(labels)} label sets for a batch of {pred.shape[0]}")
        pred = _append_labels(pred, labels, conf_thres)
    pred = pred.contiguous()
    bs, rows, no = pred.shape
    nc = no - 5
    multi_label = bool(multi_label) and nc > 1                        # :111
    buf = _cand_buffer(pred.device, torch.cuda.current_stream(pred.device).cuda_stream, bs, rows, nc, multi_label,
                       ops.expects_dense("auto", conf_thres))
    ops.filter_pred(pred, conf_thres, multi_label, classes, buf=buf)
    out = ops.nms_batched(buf, iou_thres, agnostic, max_nms, max_det, want_keep=want_keep)
    counts = out.counts.cpu().tolist()                                # the one sync
    dets = [out.dets[i, :k] for i, k in enumerate(counts)]
    if pred.dtype != torch.float32:                                   # the reference returns prediction's dtype
        dets = [d.to(pred.dtype) for d in dets]
    if want_keep:
        return dets, [out.keep[i, :k] for i, k in enumerate(counts)]
    return dets


def nms(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
        classes=None, agnostic: bool = False, multi_label: bool = False, labels=(),
        max_det: int = 300) -> List[torch.Tensor]:
    """utils/image_proc.py:83-187.  Returns a list of (k, 6) tensors
    [x1, y1, x2, y2, conf, cls] per image, descending score, on ``prediction.device``."""
    return _run_nms(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels,
                    max_det, MAX_NMS)
