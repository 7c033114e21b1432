"""Drop-in for the two hot-path helpers of ``vision_kit.utils.bboxes``."""
from __future__ import annotations

import torch

from . import ops


def cxcywh_to_xyxy(bboxes: torch.Tensor) -> torch.Tensor:
    """utils/bboxes.py:103-111, (n, 4) float32 CUDA tensor -> new tensor."""
    return ops.cxcywh_to_xyxy(bboxes)


def clip_coords(boxes: torch.Tensor, shape) -> None:
    """utils/bboxes.py:50-59, in place: a scale with gain 1 and no pad is a pure clip."""
    ops.scale_coords_(boxes, 0.0, 0.0, 1.0, False, clip_hw=shape)
