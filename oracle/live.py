"""Live import of the real reference (TEST INFRASTRUCTURE ONLY; dev container).

``/root/reference`` is read-only here and absent on the GPU box, so nothing
that runs there may import this module; callers check ``available()``.
Recipe from SURVEY.md §8c: stub the two uninstallable imports
(``pycocotools``, ``omegaconf``) and neutralise the wall-clock NMS time limit
(utils/image_proc.py:109,183-185) by freezing the clock that module sees.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("VK_REFERENCE_ROOT", "/root/reference")
_cache = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vision_kit"))


def load():
    """Returns a namespace with ``image_proc``, ``ImageProcessor``,
    ``YoloV5Head``, ``YoloV7Head`` of the unmodified reference."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    stubs = {
        "pycocotools": {},
        "pycocotools.coco": {"COCO": object},
        "omegaconf": {"OmegaConf": object, "DictConfig": dict},
        "omegaconf.dictconfig": {"DictConfig": dict},
    }
    for name, attrs in stubs.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from vision_kit.utils import image_proc
    from vision_kit.demo.processing import ImageProcessor
    from vision_kit.models.heads import YoloV5Head, YoloV7Head

    class _FrozenClock:
        @staticmethod
        def time():
            return 0.0

    image_proc.time = _FrozenClock          # time limit never fires
    ns = types.SimpleNamespace(image_proc=image_proc, ImageProcessor=ImageProcessor,
                               YoloV5Head=YoloV5Head, YoloV7Head=YoloV7Head)
    _cache["ns"] = ns
    return ns


class capture_keep:
    """Context manager recording what ``torchvision.ops.nms`` returns while the
    reference's ``nms`` runs (the reference never exposes the keep indices)."""

    def __enter__(self):
        import torchvision
        self._tv = torchvision
        self._orig = torchvision.ops.nms
        self.keeps = []

        def spy(boxes, scores, thr):
            k = self._orig(boxes, scores, thr)
            self.keeps.append(k.clone())
            return k

        torchvision.ops.nms = spy
        return self

    def __exit__(self, *exc):
        self._tv.ops.nms = self._orig
        return False


def head_decode(variant: str, levels):
    """Runs the real head's eval forward with identity convs on
    ``levels[i] = (B, 255, ny, nx)``."""
    import torch
    from torch import nn
    ns = load()
    head = ns.YoloV5Head() if variant == "v5" else ns.YoloV7Head(deploy=True)
    head.m = nn.ModuleList([nn.Identity() for _ in range(3)])
    head.eval()
    with torch.no_grad():
        pred, raws = head([t.clone() for t in levels])
    return pred, raws
