"""Drop-in for ``vision_kit.demo.processing.ImageProcessor`` (reference
demo/processing.py:11-199): same constructor, methods and stored state (``ratio``, ``pad``),
running on the sm_100a kernels.  ``preprocess`` returns the (1,3,H,W) tensor on the CUDA
device (the reference returns a CPU tensor that scripts/demo.py:70 moves with ``.to(device)``,
which is then a no-op)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .image_proc import _run_nms


class ImageProcessor:
    def __init__(self, conf_thres: float = 0.25, iou_thres: float = 0.45, filtered_classes: tuple = None,
                 labels: tuple = (), img_sz: list = (640, 640), color: list = (114, 114, 114),
                 letterbox: bool = True, auto: bool = False, scaleup: bool = True,
                 agnostic: bool = False, multi_label: bool = False, max_det: int = 300,
                 stride: int = 32) -> None:
        self.img_sz = img_sz
        self.color = color
        self.stride = stride
        self.letterbox = letterbox
        self.auto = auto
        self.scaleup = scaleup
        self.conf_thres = conf_thres
        self.iou_thres = iou_thres
        self.filtered_classes = filtered_classes
        self.agnostic = agnostic
        self.multi_label = multi_label
        self.labels = labels
        self.max_det = max_det

    # -- demo/processing.py:45-52
    def preprocess(self, img: np.ndarray, is_BGR: bool = True, make_tensor: bool = True):
        if not make_tensor:
            if is_BGR:
                img = img[:, :, ::-1]
            return self.resize(np.ascontiguousarray(img))
        src = torch.from_numpy(np.ascontiguousarray(img)).cuda()
        out, rps = ops.letterbox_batch([src], self.img_sz, self.stride, self.letterbox, self.scaleup,
                                       self.auto, self.color, swap_rb=bool(is_BGR), dtype=torch.float32)
        self._store(rps[0])
        return out, (self.ratio, self.pad)

    # -- demo/processing.py:54-57: only image 0 is un-letterboxed and returned
    def postprocess(self, prediction: torch.Tensor) -> torch.Tensor:
        outputs = self.nms(prediction)
        return self.scale_coords(outputs[0])

    # -- demo/processing.py:59-97
    def resize(self, img: np.ndarray):
        src = torch.from_numpy(np.ascontiguousarray(img)).cuda()
        out, rps = ops.letterbox_batch([src], self.img_sz, self.stride, self.letterbox, self.scaleup,
                                       self.auto, self.color, swap_rb=False, dtype=torch.uint8)
        self._store(rps[0])
        return out[0].cpu().numpy(), (self.ratio, self.pad)

    def _store(self, rp):
        ratio, pad = rp
        if not self.letterbox:
            pad = (int(pad[0]), int(pad[1]))
        self.ratio, self.pad = ratio, pad           # demo/processing.py:66,92

    # -- demo/processing.py:99-105: in place, no clip, returns the same tensor
    def scale_coords(self, outputs: torch.Tensor) -> torch.Tensor:
        return ops.scale_coords_(outputs, self.pad[0], self.pad[1], self.ratio,
                                 subtract_pad=bool(self.letterbox), clip_hw=None)

    # -- demo/processing.py:107-199 (max_nms = 10000 :119, no time limit)
    def nms(self, prediction: torch.Tensor):
        nc = prediction.shape[2] - 5
        self.multi_label &= nc > 1                   # :121 mutates state like the reference
        return _run_nms(prediction, self.conf_thres, self.iou_thres, self.filtered_classes,
                        self.agnostic, self.multi_label, self.labels, self.max_det, 10000)
