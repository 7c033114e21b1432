"""Drop-in for ``vision_kit.utils.drawing.Drawing`` (reference utils/drawing.py:108-144).

Drawing is host-side OpenCV work and stays that: what changes is the transfer -- the reference calls
``det.cpu().numpy()`` once per detection (a device synchronisation each, scripts/demo.py:76 hands it the
CUDA tensor of ``postprocess``); here the whole (k, 6) tensor comes over in one copy.  Geometry, label
text, font and the filled / outlined modes are the reference's.  The reference's 140-entry colour table is
its own data and is not duplicated here: colours come from ``palette`` (any ``index -> (b, g, r)``
callable -- pass the reference's ``COLOR("bgr")`` for identical pixels) or from a generated table.
"""
from __future__ import annotations

import colorsys
from typing import Callable, Optional, Sequence

import cv2
import numpy as np
import torch


def _generated_palette(n: int = 140):
    cols = []
    for i in range(n):
        r, g, b = colorsys.hsv_to_rgb((i * 0.61803398875) % 1.0, 0.55 + 0.45 * ((i * 7) % 3) / 2.0, 1.0 - 0.35 * ((i * 5) % 4) / 3.0)
        cols.append((int(b * 255), int(g * 255), int(r * 255)))
    return cols


class Drawing:
    def __init__(self, class_names: Sequence[str], palette: Optional[Callable[[int], tuple]] = None) -> None:
        self._class_names = class_names
        table = _generated_palette()
        self._color = palette if palette is not None else (lambda i: table[int(i) % len(table)])

    def draw(self, img: np.ndarray, dets, filled: bool = False) -> np.ndarray:
        """``dets``: (k, 6) tensor / array or a list of (6,) tensors [x1, y1, x2, y2, score, cls]."""
        if torch.is_tensor(dets):
            rows = dets.detach().cpu().numpy()                        # one transfer for every box
        elif len(dets) and torch.is_tensor(dets[0]):
            rows = torch.stack(list(dets)).detach().cpu().numpy()
        else:
            rows = np.asarray(dets, np.float32).reshape(-1, 6)
        font = cv2.FONT_HERSHEY_SIMPLEX
        for pred in rows:
            x0, y0, x1, y1 = map(int, pred[:4].tolist())              # drawing.py:117-118
            label = int(pred[-1])
            score = pred[-2].item()
            color = self._color(label)
            text = "{}:{:.1f}%".format(self._class_names[label], score * 100)
            txt_color = (0, 0, 0) if (np.mean(color) / 255) > 0.5 else (255, 255, 255)
            txt_size = cv2.getTextSize(text, font, 0.4, 1)[0]
            if filled:                                                # :127-133
                overlay = img.copy()
                cv2.rectangle(overlay, (x0, y0), (x1, y1), color, -1)
                img = cv2.addWeighted(overlay, 0.5, img, 0.5, 0)
            else:
                cv2.rectangle(img, (x0, y0), (x1, y1), color, 2)
            cv2.rectangle(img, (x0, y0 + 1), (x0 + txt_size[0] + 1, y0 + int(1.5 * txt_size[1])), color, -1)
            cv2.putText(img, text, (x0, y0 + txt_size[1]), font, 0.4, txt_color, 1)
        return img
