"""Drop-in for the matching half of ``vision_kit.core.eval.det_evaluator.DetEvaluator``
(core/eval/det_evaluator.py:100-182, 274-300): ``evaluate`` and ``process_batch`` with the
reference's signatures, the per-image loop replaced by ONE kernel launch for the batch
(``vk_eval_match``: un-letterbox + clip, torchvision box_iou, greedy unique matching at the ten
IoU thresholds).  ``stats`` holds the same four tensors per image as the reference's, so the
reference's own ``summarize`` / ``ap_per_class`` (host-side NumPy, out of scope here) consume it
unchanged.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib, ops


class DetEvaluator:
    def __init__(self, class_labels: list, img_size: Tuple[int, int] = (640, 640), gt_json: str = None,
                 label_format: str = "yolo") -> None:
        self.class_labels = class_labels
        self.img_sz = img_size
        self.class_ids = [i + 1 for i in range(len(self.class_labels))]
        self.gt_json = gt_json
        self.label_format = label_format
        self.iouv = torch.linspace(0.5, 0.95, 10)          # :123 (float32)
        self.num_iou = self.iouv.numel()
        self.seen = 0
        self.stats: list = []

    # ------------------------------------------------------------------ batch entry
    def evaluate(self, img: torch.Tensor, img_infos: Sequence, idxs: Sequence, preds, targets: torch.Tensor):
        """core/eval/det_evaluator.py:129-182.  ``preds``: the reference's list of (k, 6) tensors
        or an ``ops.NmsOut`` (padded (B, max_det, 6) + counts, no packing needed).  Like the
        reference, ``targets[:, 2:]`` is scaled to pixels IN PLACE for label_format 'yolo'.
        Returns (vstack of predn, vstack of targetn) over the images that have detections."""
        _lib.require_cuda(targets, "targets")
        dev = targets.device
        self.iouv = self.iouv.to(dev)
        b, _, h, w = img.shape
        if self.label_format == "yolo":
            targets[:, 2:] *= torch.tensor((w, h, w, h), device=dev)                       # :138-139
        if isinstance(preds, ops.NmsOut):
            dets, counts = preds.dets, preds.counts
        else:
            max_det = max([int(p.shape[0]) for p in preds] + [1])
            dets = torch.zeros((b, max_det, 6), dtype=torch.float32, device=dev)
            for i, p in enumerate(preds):
                dets[i, : p.shape[0]] = p
            counts = torch.tensor([int(p.shape[0]) for p in preds], dtype=torch.int32, device=dev)
        # group the labels by image (the reference selects `targets[:, 0] == idx` per image, :143)
        img_col = targets[:, 0].long()
        order = torch.argsort(img_col, stable=True)
        labels = targets[order].contiguous()
        per_img = torch.bincount(img_col, minlength=b)[:b]
        offsets = torch.zeros(b + 1, dtype=torch.int32, device=dev)
        offsets[1:] = torch.cumsum(per_img, 0)
        n_lbl = per_img.tolist()                                                           # one host sync per batch
        img0 = torch.tensor([[int(s[0]), int(s[1])] for s in img_infos], dtype=torch.int32, device=dev)
        m = ops.eval_match(dets, counts, labels, offsets, max(n_lbl + [0]), img0, (h, w), self.iouv)
        n_pred = counts.tolist()
        predictions, detections = [], []
        lo = 0
        for i in range(b):
            self.seen += 1
            k, nl = n_pred[i], n_lbl[i]
            tcls = labels[lo: lo + nl, 1]
            targetn = m.labeln[lo: lo + nl]
            lo += nl
            if k == 0:
                if nl:                                                                     # :157-162
                    self.stats.append((torch.zeros(0, self.num_iou, dtype=torch.bool, device=dev),
                                       torch.zeros(0, device=dev), torch.zeros(0, device=dev), tcls))
                continue
            self.stats.append((m.correct[i, :k], dets[i, :k, 4], dets[i, :k, 5], tcls))    # :172
            predictions.append(m.predn[i, :k])
            detections.append(targetn)
        if not predictions:
            return torch.zeros((0, 6), device=dev), torch.zeros((0, 5), device=dev)
        return torch.vstack(predictions), torch.vstack(detections)

    # ------------------------------------------------------------------ single image
    @staticmethod
    def process_batch(preds: torch.Tensor, labels: torch.Tensor, iouv: torch.Tensor) -> torch.Tensor:
        """:274-300.  preds (N, 6) x1,y1,x2,y2,conf,cls and labels (M, 5) cls,x1,y1,x2,y2 already in
        the same pixel frame -> bool (N, len(iouv)).  The batch kernel on one image, nothing rescaled."""
        _lib.require_cuda(preds, "preds")
        dev = preds.device
        n, m = int(preds.shape[0]), int(labels.shape[0])
        if n == 0:
            return torch.zeros((0, int(iouv.numel())), dtype=torch.bool, device=dev)
        lab6 = torch.cat([torch.zeros((m, 1), device=dev), labels.float()], 1)        # image 0, cls, x1, y1, x2, y2
        out = ops.eval_match(preds.float().contiguous().view(1, n, 6), torch.tensor([n], dtype=torch.int32, device=dev),
                             lab6, torch.tensor([0, m], dtype=torch.int32, device=dev), m, None, (1, 1),
                             iouv.to(dev), prescaled=True)
        return out.correct[0, :n]
