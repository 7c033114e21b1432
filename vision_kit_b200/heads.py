"""Drop-in for ``vision_kit.models.heads`` (reference models/heads/yolov5.py,
models/heads/yolov7.py): same constructors, attributes and return structure.  The 1x1
convs (and v7's ImplicitA/M) stay PyTorch modules; everything after them in eval mode --
view/permute, sigmoid, grid and anchor decode, concat -- is one launch of the
``detect_decode`` kernel, and ``forward_nms`` skips the (B, rows, no) tensor altogether with
the fused decode+filter kernel."""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn

from . import ops


def check_anchor_order(anchors: torch.Tensor, stride: torch.Tensor) -> torch.Tensor:
    """utils/model_utils.py:72-81."""
    a = anchors.prod(-1).mean(-1).view(-1)
    da = a[-1] - a[0]
    ds = stride[-1] - stride[0]
    if da and (da.sign() != ds.sign()):
        anchors[:] = anchors.flip(0)
    return anchors


def init_bias(module: nn.ModuleList, stride, na: int, nc: int, cf=None):
    """utils/model_utils.py:37-43."""
    for m, s in zip(module, stride):
        b = m.bias.view(na, -1)
        b.data[:, 4] += math.log(8 / (640 / float(s)) ** 2)
        b.data[:, 5:] += math.log(0.6 / (nc - 0.99)) if cf is None else torch.log(cf / cf.sum())
        m.bias = torch.nn.Parameter(b.view(-1), requires_grad=True)


class Implicit(nn.Module):
    """models/modules/blocks.py:494-517."""

    def __init__(self, channel: int, ops: str = "add", mean: float = None, std: float = .02) -> None:
        super().__init__()
        assert ops.lower() in ["add", "multiply"], "Not Implemented Operation!"
        self.channel = channel
        self.ops = ops.lower()
        self.mean = mean if mean else 0.0 if self.ops == "add" else 1.0
        self.std = std
        weight = torch.zeros(1, channel, 1, 1) if self.ops == "add" else torch.ones(1, channel, 1, 1)
        self.implicit = nn.Parameter(weight)
        nn.init.normal_(self.implicit, mean=self.mean, std=self.std)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.implicit + x if self.ops == "add" else self.implicit * x


class _DetectBase(nn.Module):
    variant = "v5"

    def _conv(self, i: int, x: torch.Tensor) -> torch.Tensor:
        return self.m[i](x)

    def _anchors_px(self) -> List[List[float]]:
        raise NotImplementedError

    def _cfg(self, feats):
        grids = [(int(f.shape[2]), int(f.shape[3])) for f in feats]
        key = tuple(grids)
        if getattr(self, "_cfg_key", None) != key:
            self._cfg_cache = ops.head_cfg(self.variant, self.num_classes, self._anchors_px(),
                                           [float(s) for s in self.stride], grids)
            self._cfg_key = key
        return self._cfg_cache

    def _conv_all(self, x):
        # the kernels read float32, float16 and bfloat16 conv outputs as they are (AMP eval emits
        # float16, scripts/main.py:41): no up-cast pass
        return [self._conv(i, x[i]).contiguous() for i in range(self.num_det_layers)]

    def forward(self, x):
        x = list(x)
        if self.training:                # heads/yolov5.py:57-60,78: raw permuted maps only
            for i in range(self.num_det_layers):
                x[i] = self._conv(i, x[i])
                bs, _, ny, nx = x[i].shape
                x[i] = x[i].view(bs, self.num_anchors, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            return x
        feats = self._conv_all(x)
        cfg = self._cfg(feats)
        if self.export:                  # :78 -> (pred,)
            return (ops.detect_decode(cfg, feats, want_raw=False),)
        pred, raws = ops.detect_decode(cfg, feats, want_raw=True)
        return pred, raws

    def _fused_conv_params(self):
        """(weights, biases) of the 1x1 convs as plain matrices, implicit layers folded in
        (heads/yolov7.py:67-71: im * (W (x + ia) + b) = (im * W) x + im * (W ia + b))."""
        ws, bs = [], []
        for i in range(self.num_det_layers):
            w = self.m[i].weight.detach().reshape(self.m[i].out_channels, -1).float()
            b = self.m[i].bias.detach().float() if self.m[i].bias is not None else torch.zeros(w.shape[0], device=w.device)
            if hasattr(self, "ia"):
                ia = self.ia[i].implicit.detach().reshape(-1).float()
                im = self.im[i].implicit.detach().reshape(-1).float()
                b = im * (w @ ia + b)
                w = im[:, None] * w
            ws.append(w.contiguous())
            bs.append(b.contiguous())
        return ws, bs

    @torch.no_grad()
    def forward_nms(self, x, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                    agnostic: bool = False, multi_label: bool = False, max_det: int = 300,
                    max_nms: int = 30000, fused_conv: bool = False):
        """Fused eval path: conv outputs -> candidates -> NMS, no prediction tensor.
        ``fused_conv=True`` also folds the Detect 1x1 convs in (``vk_conv_decode_filter``: tcgen05,
        TF32 inputs -- what cuDNN does by default for fp32 convs -- fp32 accumulation): the neck
        outputs go in, the (B, 255, ny, nx) conv outputs are never written.
        Returns an ``ops.NmsOut`` (device tensors; no synchronisation).  With ``fused_conv`` its ``fault``
        member carries the tensor-core kernel's time-out flag; ``NmsOut.check()`` / ``DetectPipeline.to_list``
        raise if it is set."""
        if fused_conv:
            feats = [f.float().contiguous() for f in x]
            cfg = self._cfg(feats)
            ws, bs = self._fused_conv_params()
            buf = ops.conv_decode_filter(cfg, feats, ws, bs, conf_thres, multi_label, classes)
        else:
            feats = self._conv_all(list(x))
            cfg = self._cfg(feats)
            buf = ops.decode_filter(cfg, feats, conf_thres, multi_label, classes)
        return ops.nms_batched(buf, iou_thres, agnostic, max_nms, max_det)


class YoloV5Head(_DetectBase):
    variant = "v5"

    def __init__(self, num_classes: int = 80, width: float = 1.00, anchors: list = None,
                 in_chs: tuple = (256, 512, 1024), stride: list = [8., 16., 32.],
                 deploy: bool = False, export: bool = False) -> None:
        super().__init__()
        if anchors is None:
            anchors = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119],
                       [116, 90, 156, 198, 373, 326]]                     # yolov5.py:24-28
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.num_classes = num_classes
        self.no = num_classes + 5
        self.num_det_layers = len(anchors)
        self.num_anchors = len(anchors[0]) // 2
        self.stride = torch.tensor(stride, device=self.device)
        self.anchors = torch.tensor(anchors, device=self.device).float().view(self.num_det_layers, -1, 2)
        self.anchors /= self.stride.view(-1, 1, 1)                        # :41-43
        self.anchors = check_anchor_order(self.anchors, self.stride)
        self.m = nn.ModuleList(nn.Conv2d(int(x * width), self.no * self.num_anchors, 1) for x in in_chs)
        self.export = export
        init_bias(self.m, self.stride, self.num_anchors, self.num_classes)

    def _anchors_px(self):
        # anchor_grid = anchors[i] * stride[i]  (yolov5.py:89)
        return [(self.anchors[i] * self.stride[i]).reshape(-1).tolist() for i in range(self.num_det_layers)]


class YoloV7Head(_DetectBase):
    variant = "v7"

    def __init__(self, variant: str = "base", num_classes: int = 80, anchors: list = None,
                 stride: tuple = (8., 16., 32.), deploy: bool = False, export: bool = False) -> None:
        super().__init__()
        if anchors is None:
            anchors = [[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146],
                       [142, 110, 192, 243, 459, 401]]                    # yolov7.py:23-27
        in_chs = {"base": [256, 512, 1024], "x": [320, 640, 1280]}[variant.lower()]
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.export = export
        self.deploy = deploy
        self.num_classes = num_classes
        self.no = num_classes + 5
        self.num_det_layers = len(anchors)
        self.num_anchors = len(anchors[0]) // 2
        self.stride = torch.tensor(stride, device=self.device)
        self.anchors = torch.tensor(anchors, device=self.device).float().view(self.num_det_layers, -1, 2)
        self.anchor_grid = self.anchors.clone().view(self.num_det_layers, 1, -1, 1, 1, 2)   # :47
        self.anchors /= self.stride.view(-1, 1, 1)
        self.anchors = check_anchor_order(self.anchors, self.stride)
        self.m = nn.ModuleList(nn.Conv2d(x, self.no * self.num_anchors, 1) for x in in_chs)
        if not self.deploy:
            self.ia = nn.ModuleList(Implicit(x, ops="add") for x in in_chs)
            self.im = nn.ModuleList(Implicit(self.no * self.num_anchors, ops="multiply") for _ in in_chs)
        init_bias(self.m, self.stride, self.num_anchors, self.num_classes)

    def _conv(self, i: int, x: torch.Tensor) -> torch.Tensor:
        if self.training or hasattr(self, "ia"):                          # yolov7.py:67-71
            return self.im[i](self.m[i](self.ia[i](x)))
        return self.m[i](x)

    def _anchors_px(self):
        return [self.anchor_grid[i].reshape(-1).tolist() for i in range(self.num_det_layers)]   # :81
