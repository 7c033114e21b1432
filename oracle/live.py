"""Live import of the real reference (TEST INFRASTRUCTURE ONLY).

``/root/reference`` exists in the dev container only.  ``__graft_entry__.build()`` therefore installs the
unmodified reference package into ``baseline/_ref`` (git-ignored; it travels to the GPU box with the
snapshot) with ``pip install --no-deps --target``, and this module imports it from whichever of the two
is there; callers check ``available()``.
Recipe from SURVEY.md §8c: stub the two uninstallable imports
(``pycocotools``, ``omegaconf``) and neutralise the wall-clock NMS time limit
(utils/image_proc.py:109,183-185) by freezing the clock that module sees.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("VK_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "baseline", "_ref")]
REF_ROOT = next((p for p in _CANDIDATES if p and os.path.isdir(os.path.join(p, "vision_kit"))), "/root/reference")
ASSETS = next((p for p in (os.path.join(REF_ROOT, "assets"), os.path.join(_HERE, "baseline", "_ref", "assets"))
               if os.path.isdir(p)), None)
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
                               YoloV5Head=YoloV5Head, YoloV7Head=YoloV7Head, root=REF_ROOT)
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


def demo_model(seed: int = 0):
    """The demo's model (scripts/demo.py:48-53: ``YOLOV5(variant='s')``) with seeded random weights -- the
    weight file is not part of the reference tree -- conditioned so that it produces a few hundred
    detections on a photograph: BatchNorm layers use batch statistics (with the initial running statistics
    the activations of a random network vanish and every logit equals its bias), the Detect convolutions are
    scaled by 4 and their objectness / class priors lifted.  Modules and code paths are the reference's own.
    One model per process: the neck mutates its default arguments."""
    import torch
    load()
    from vision_kit.models.architectures import YOLOV5
    torch.manual_seed(seed)
    model = YOLOV5("s").eval()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.train()
    with torch.no_grad():
        for m in model.head.m:
            m.weight.mul_(4.0)
            b = m.bias.view(3, -1)
            b[:, 4] += 0.3
            b[:, 5:] += 2.5
    return model
