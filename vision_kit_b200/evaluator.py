"""Drop-in for the matching half of ``vision_kit.core.eval.det_evaluator.DetEvaluator``
(core/eval/det_evaluator.py:100-182, 274-300): ``evaluate`` and ``process_batch`` with the
reference's signatures, the per-image loop replaced by ONE kernel launch for the batch
(``vk_eval_match``: un-letterbox + clip, torchvision box_iou, greedy unique matching at the ten
IoU thresholds).  ``stats`` holds the same four tensors per image as the reference's, so the
reference's own ``summarize`` / ``ap_per_class`` consume it unchanged; ``summarize`` here restates
that host-side NumPy arithmetic (:13-97, 184-195; SURVEY.md §8f row 4 -- float64, no kernel) so the
class is usable on its own.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
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
        self.coco_data: list = []
        self.precision = self.recall = self.f1 = 0.0
        self.mp = self.mr = self.map50 = self.map95 = 0.0

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

    # ------------------------------------------------------------------ COCO json export
    def convert_to_coco(self, pred: torch.Tensor, img_id) -> None:
        """core/eval/det_evaluator.py:228-244: detections (k, 6) of one image, already in original-image
        pixels, appended to ``coco_data`` as COCO result dicts (bbox = x, y, w, h; category = class id + 1)."""
        self._append_coco(pred.detach().cpu().numpy(), img_id)

    def convert_batch_to_coco(self, predn: torch.Tensor, counts, img_ids: Sequence) -> None:
        """The same for a whole batch in ONE device-to-host copy: ``predn`` (B, max_det, 6) padded (the
        ``predn`` of ``ops.eval_match`` / an ``NmsOut.dets`` in original pixels) + counts."""
        host = predn.detach().cpu().numpy()
        ks = counts.tolist() if torch.is_tensor(counts) else list(counts)
        for i, k in enumerate(ks):
            self._append_coco(host[i, : int(k)], img_ids[i])

    def _append_coco(self, p: np.ndarray, img_id) -> None:
        p = np.asarray(p, np.float32)
        if p.size == 0:
            return
        xywh = p[:, :4].copy()
        xywh[:, 2] = p[:, 2] - p[:, 0]                       # utils/bboxes.py:114-119 xyxy_to_xywh
        xywh[:, 3] = p[:, 3] - p[:, 1]
        for row, score, c in zip(xywh.tolist(), p[:, 4], p[:, 5]):
            self.coco_data.append({"image_id": int(img_id), "category_id": self.class_ids[int(c)], "bbox": row,
                                   "score": score.item(), "segmentation": []})

    def coco_evaluate(self) -> str:
        """:246-272, verbatim in behaviour: needs pycocotools and ``gt_json`` (neither is part of this
        repo's scope; the import error of a missing pycocotools is the reference's own)."""
        import contextlib
        import io
        import json
        import tempfile

        from pycocotools.coco import COCO
        from pycocotools.cocoeval import COCOeval
        info = ""
        if len(self.coco_data) > 0:
            gt = COCO(self.gt_json)
            _, tmp = tempfile.mkstemp()
            with open(tmp, "w") as f:
                json.dump(self.coco_data, f)
            dt = gt.loadRes(tmp)
            ev = COCOeval(gt, dt, "bbox")
            ev.evaluate()
            ev.accumulate()
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                ev.summarize()
            info = buf.getvalue()
        return info

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


# ---------------------------------------------------------------------------------------------
# AP accumulation (host, float64): core/eval/det_evaluator.py:13-97 + utils/metrics.py:15-20
# ---------------------------------------------------------------------------------------------
_PR_GRID = np.linspace(0, 1, 1000)      # confidence grid of the P/R/F1 curves (:35)
_AP_GRID = np.linspace(0, 1, 101)       # 101-point COCO interpolation (:87)


def _envelope_ap(recall: np.ndarray, precision: np.ndarray) -> float:
    """:70-97 ('interp' method): sentinels, monotone precision envelope from the right, trapezoid
    over 101 recall points."""
    r = np.concatenate(([0.0], recall, [1.0]))
    p = np.concatenate(([1.0], precision, [0.0]))
    p = np.maximum.accumulate(p[::-1])[::-1]
    y = np.interp(_AP_GRID, r, p)
    return float(np.sum((y[1:] + y[:-1]) * np.diff(_AP_GRID) / 2.0))     # == np.trapz(y, x)


def _box_smooth(y: np.ndarray, frac: float) -> np.ndarray:
    """utils/metrics.py:15-20: box filter over a fraction of the curve, edge-padded."""
    n = round(len(y) * frac * 2) // 2 + 1
    half = np.ones(n // 2)
    return np.convolve(np.concatenate((half * y[0], y, half * y[-1])), np.ones(n) / n, mode="valid")


def average_precision_per_class(tp, conf, pred_cls, target_cls, eps: float = 1e-16):
    """``ap_per_class`` (:13-67).  tp (n, niou) bool/0-1, conf (n,), pred_cls (n,), target_cls (m,).
    Returns (tp, fp, precision, recall, f1, ap (nc, niou), classes) at the max-F1 confidence."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, n_labels = np.unique(target_cls, return_counts=True)
    niou = tp.shape[1]
    ap = np.zeros((len(classes), niou))
    p_curve = np.zeros((len(classes), _PR_GRID.size))
    r_curve = np.zeros_like(p_curve)
    for k, (c, nl) in enumerate(zip(classes, n_labels)):
        sel = pred_cls == c
        if not sel.any() or nl == 0:
            continue
        hits = tp[sel].cumsum(0)
        misses = (1 - tp[sel]).cumsum(0)
        recall = hits / (nl + eps)
        precision = hits / (hits + misses)
        # curves over confidence at IoU 0.5; -x because np.interp wants increasing abscissae
        r_curve[k] = np.interp(-_PR_GRID, -conf[sel], recall[:, 0], left=0)
        p_curve[k] = np.interp(-_PR_GRID, -conf[sel], precision[:, 0], left=1)
        for j in range(niou):
            ap[k, j] = _envelope_ap(recall[:, j], precision[:, j])
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    best = _box_smooth(f1_curve.mean(0), 0.1).argmax()
    p, r, f1 = p_curve[:, best], r_curve[:, best], f1_curve[:, best]
    tp_n = (r * n_labels).round()
    fp_n = (tp_n / (p + eps) - tp_n).round()
    return tp_n, fp_n, p, r, f1, ap, classes.astype(int)


def _summarize(self, details_per_class: bool = False, do_coco_eval: bool = False):
    """``DetEvaluator.summarize`` (:184-226) without the rich table / pycocotools parts: returns
    (map50, map95, per_class rows or None, None) and resets the accumulators."""
    stats = [torch.cat(x, 0).cpu().numpy() for x in zip(*self.stats)] if self.stats else []
    rows = None
    ap_class = np.zeros(0, int)
    ap50 = ap = np.zeros(0)
    if len(stats) and stats[0].any():
        _, _, self.precision, self.recall, self.f1, ap_all, ap_class = average_precision_per_class(*stats)
        ap50, ap = ap_all[:, 0], ap_all.mean(1)
        self.mp, self.mr = self.precision.mean(), self.recall.mean()
        self.map50, self.map95 = ap50.mean(), ap.mean()
    if details_per_class and len(stats):
        counts = np.bincount(stats[3].astype(int), minlength=len(self.class_labels))
        rows = [[self.class_labels[int(c)], self.seen, int(counts[c]), round(float(self.precision[i]), 3),
                 round(float(self.recall[i]), 3), round(float(ap50[i]), 3), round(float(ap[i]), 3)]
                for i, c in enumerate(ap_class)]
    self.seen = 0
    self.stats.clear()
    return self.map50, self.map95, rows, None


DetEvaluator.summarize = _summarize
