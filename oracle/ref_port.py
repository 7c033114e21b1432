"""Library-level port of the reference CPU path (TEST INFRASTRUCTURE ONLY).

The reference is pure Python over cv2 / torch / torchvision; this module states
the same path against the same libraries so that it can (a) run on the GPU
box, where ``/root/reference`` does not exist, as the CPU baseline that
``bench.py`` times (``cpu_baseline.kind = "port"`` and ``--impl reference``),
and (b) serve as a second oracle next to the numpy restatement.  It is pinned
against the live reference by tests/test_oracle_golden.py.

Reference lines followed (relative to /root/reference/vision_kit/):
  letterbox      utils/image_proc.py:12-60, demo/processing.py:59-97
  preprocess     demo/processing.py:45-52
  detect_decode  models/heads/yolov5.py:54-91, models/heads/yolov7.py:62-95
  nms            utils/image_proc.py:83-187, demo/processing.py:107-199
  scale_coords   utils/image_proc.py:63-80, demo/processing.py:99-105
"""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torchvision

from .restate import letterbox_geometry


def letterbox(img: np.ndarray, img_sz=(640, 640), stride=32, letterbox=True,
              scaleup=True, auto=False, color=(114, 114, 114)):
    g = letterbox_geometry(img.shape[0], img.shape[1], img_sz, stride,
                           letterbox, scaleup, auto)
    if (img.shape[1], img.shape[0]) != (g["new_w"], g["new_h"]):
        img = cv2.resize(img, (g["new_w"], g["new_h"]),
                         interpolation=cv2.INTER_LINEAR)
    img = cv2.copyMakeBorder(img, g["top"], g["bottom"], g["left"], g["right"],
                             cv2.BORDER_CONSTANT, value=color)
    return img, (g["ratio"], g["pad"])


def preprocess(img: np.ndarray, img_sz=(640, 640), is_bgr=True, **kw):
    if is_bgr:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    img, rp = letterbox(img, img_sz, **kw)
    chw = np.ascontiguousarray(img.transpose(2, 0, 1))
    return torch.from_numpy(chw).unsqueeze(0) / 255, rp


def detect_decode(levels, anchors_px, strides, variant: str):
    """``levels[i]``: float32 tensor (B, na*no, ny, nx), the 1x1-conv output."""
    preds, raws = [], []
    for i, x in enumerate(levels):
        anc = torch.as_tensor(anchors_px[i], dtype=torch.float32).view(-1, 2)
        na = anc.shape[0]
        bs, ch, ny, nx = x.shape
        no = ch // na
        raw = x.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raws.append(raw)
        y = raw.sigmoid()
        gy, gx = torch.meshgrid(torch.arange(ny, dtype=torch.float32),
                                torch.arange(nx, dtype=torch.float32), indexing="ij")
        grid = torch.stack((gx, gy), 2).view(1, 1, ny, nx, 2)
        if variant == "v5":
            y[..., 0:2] = (y[..., 0:2] * 2 + (grid - 0.5)) * float(strides[i])
        else:
            y[..., 0:2] = (y[..., 0:2] * 2. - 0.5 + grid) * float(strides[i])
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anc.view(1, na, 1, 1, 2)
        preds.append(y.view(bs, -1, no))
    return torch.cat(preds, 1), raws


def nms(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None,
        agnostic=False, multi_label=False, max_det=300, max_nms=30000,
        max_wh=7680, return_keep=False):
    """Per-image filter + ``torchvision.ops.nms`` with the class-offset trick.
    No wall-clock time limit; the ``n > max_nms`` cut uses a stable argsort
    (SURVEY.md §7)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    nc = prediction.shape[2] - 5
    multi_label = multi_label and nc > 1
    outs, keeps = [], []
    for x in prediction:
        x = x[x[:, 4] > conf_thres]
        if x.shape[0] == 0:
            outs.append(torch.zeros((0, 6)))
            keeps.append(torch.zeros((0,), dtype=torch.int64))
            continue
        x[:, 5:] *= x[:, 4:5]
        cx, cy, hw, hh = x[:, 0], x[:, 1], x[:, 2] / 2, x[:, 3] / 2
        box = torch.stack((cx - hw, cy - hh, cx + hw, cy + hh), 1)
        if multi_label:
            r, c = (x[:, 5:] > conf_thres).nonzero(as_tuple=True)
            d = torch.cat((box[r], x[r, c + 5, None], c[:, None].float()), 1)
        else:
            conf, c = x[:, 5:].max(1, keepdim=True)
            d = torch.cat((box, conf, c.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            d = d[(d[:, 5:6] == torch.tensor(classes)).any(1)]
        if d.shape[0] == 0:
            outs.append(torch.zeros((0, 6)))
            keeps.append(torch.zeros((0,), dtype=torch.int64))
            continue
        if d.shape[0] > max_nms:
            d = d[d[:, 4].argsort(descending=True, stable=True)[:max_nms]]
        off = d[:, 5:6] * (0 if agnostic else max_wh)
        k = torchvision.ops.nms(d[:, :4] + off, d[:, 4], iou_thres)[:max_det]
        outs.append(d[k])
        keeps.append(k)
    return (outs, keeps) if return_keep else outs


def scale_coords(img1_shape, coords: torch.Tensor, img0_shape, ratio_pad=None):
    before = coords.clone()
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = ((img1_shape[1] - img0_shape[1] * gain) / 2,
               (img1_shape[0] - img0_shape[0] * gain) / 2)
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, 0].clamp_(0, img0_shape[1])
    coords[:, 1].clamp_(0, img0_shape[0])
    coords[:, 2].clamp_(0, img0_shape[1])
    coords[:, 3].clamp_(0, img0_shape[0])
    return before
