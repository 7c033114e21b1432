"""Seeded synthetic workloads for tests and bench.py.

There is no network for datasets or checkpoints, so every input of the hot
path is synthetic.  All generators use numpy's PCG64 bit stream and IEEE
add / mul only (no exp / log / Box-Muller), so the same seed gives the same
bits on any host -- the golden fixtures under tests/golden depend on that.
"Normal" draws are Irwin-Hall(3) sums scaled to unit variance.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
STRIDES = (8, 16, 32)
V5_ANCHORS = ((10, 13, 16, 30, 33, 23), (30, 61, 62, 45, 59, 119),
              (116, 90, 156, 198, 373, 326))
V7_ANCHORS = ((12, 16, 19, 36, 40, 28), (36, 75, 76, 55, 72, 146),
              (142, 110, 192, 243, 459, 401))


def _normalish(rng, shape):
    u = rng.random(shape, dtype=F32)
    u += rng.random(shape, dtype=F32)
    u += rng.random(shape, dtype=F32)
    u -= F32(1.5)
    u *= F32(2.0)
    return u


def images_u8(batch: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """(B, h, w, 3) uint8 noise -- worst case for a bilinear kernel's parity."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(batch, h, w, 3), dtype=np.uint8)


def image_u8(h: int, w: int, seed: int = 0) -> np.ndarray:
    return images_u8(1, h, w, seed)[0]


def head_logits(batch: int, seed: int = 0, nc: int = 80, img: int = 640,
                na: int = 3, clusters: int = 0, obj_mean: float = -6.0):
    """Detect-head conv outputs: list of float32 (B, na*(5+nc), ny, nx) for
    strides 8/16/32.  box ~ N(0,1), obj ~ N(obj_mean,2), cls ~ N(-4,1.5).
    ``clusters`` > 0 plants that many objects per image, each replicated with
    jitter over a 3x3 cell neighbourhood and all anchors of one level with
    obj logit +4 and one hot class, so that NMS has something to suppress.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    no = 5 + nc
    out = []
    for s in STRIDES:
        n = img // s
        x = _normalish(rng, (batch, na, no, n, n))
        x[:, :, 4] = x[:, :, 4] * F32(2.0) + F32(obj_mean)
        x[:, :, 5:] = x[:, :, 5:] * F32(1.5) + F32(-4.0)
        out.append(x)
    if clusters:
        for b in range(batch):
            for _ in range(clusters):
                lvl = int(rng.integers(0, 3))
                n = img // STRIDES[lvl]
                gx, gy = int(rng.integers(0, n)), int(rng.integers(0, n))
                cls = int(rng.integers(0, nc))
                lw, lh = (rng.random(2, dtype=F32) * F32(2) - F32(1))
                x = out[lvl]
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        cx, cy = gx + dx, gy + dy
                        if not (0 <= cx < n and 0 <= cy < n):
                            continue
                        for a in range(na):
                            j = rng.random(6, dtype=F32) - F32(0.5)
                            x[b, a, 0, cy, cx] = j[0] * F32(2)
                            x[b, a, 1, cy, cx] = j[1] * F32(2)
                            x[b, a, 2, cy, cx] = lw + j[2] * F32(0.3)
                            x[b, a, 3, cy, cx] = lh + j[3] * F32(0.3)
                            x[b, a, 4, cy, cx] = F32(4) + j[4] * F32(4)
                            x[b, a, 5 + cls, cy, cx] = F32(3) + j[5] * F32(4)
    return [np.ascontiguousarray(x.reshape(batch, na * no, x.shape[3], x.shape[4]))
            for x in out]


def prediction(batch: int, rows: int = 25200, nc: int = 80, seed: int = 0,
               mode: str = "demo", clusters: int = 0, img: int = 640) -> np.ndarray:
    """A decoded prediction tensor (B, rows, 5+nc) float32 built WITHOUT
    transcendental functions so that it is bit-reproducible: cx,cy uniform on
    the canvas, w,h = 6 + 250*u^2, obj and class probabilities are powers of
    uniforms.  ``mode='demo'`` gives ~1% rows above 0.25; ``mode='eval'`` gives
    ~260 k (row, class) pairs above 0.001 per 25200 rows.
    ``clusters`` plants groups of heavily overlapping high-score rows.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.empty((batch, rows, 5 + nc), F32)
    p[..., 0:2] = rng.random((batch, rows, 2), dtype=F32) * F32(img)
    u = rng.random((batch, rows, 2), dtype=F32)
    p[..., 2:4] = F32(6) + F32(250) * u * u
    u = rng.random((batch, rows), dtype=F32)
    if mode == "demo":
        u2 = u * u
        p[..., 4] = u2 * u2 * u2 * u2            # u^8: P(obj > .25) ~ 16%
        p[..., 4] *= rng.random((batch, rows), dtype=F32) < F32(0.06)
        c = rng.random((batch, rows, nc), dtype=F32)
        c *= c
        c *= c
        c *= F32(0.6)                            # planted clusters outrank noise
        p[..., 5:] = c
    else:
        p[..., 4] = u * u                        # ~97% rows above 0.001
        c = rng.random((batch, rows, nc), dtype=F32)
        for _ in range(5):
            c *= c                               # u^32
        c *= F32(0.6)
        p[..., 5:] = c
    if clusters:
        for b in range(batch):
            for _ in range(clusters):
                k = int(rng.integers(4, 28))
                idx = rng.integers(0, rows, size=k)
                cx, cy = rng.random(2, dtype=F32) * F32(img)
                w, h = F32(20) + rng.random(2, dtype=F32) * F32(200)
                cls = int(rng.integers(0, nc))
                j = rng.random((k, 6), dtype=F32) - F32(0.5)
                p[b, idx, 0] = cx + j[:, 0] * w * F32(0.25)
                p[b, idx, 1] = cy + j[:, 1] * h * F32(0.25)
                p[b, idx, 2] = w * (F32(1) + j[:, 2] * F32(0.3))
                p[b, idx, 3] = h * (F32(1) + j[:, 3] * F32(0.3))
                p[b, idx, 4] = F32(0.9) + j[:, 4] * F32(0.2)
                p[b, idx, 5 + cls] = F32(0.9) + j[:, 5] * F32(0.2)
    return p


def mixed_sizes(batch: int, seed: int = 0, lo: int = 480, hi: int = 1280):
    """(h, w) pairs for BASELINE config 5: uniform in [lo, hi], with the
    exact-2x, up-scale and non-square corner cases forced in first."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fixed = [(1280, 1280), (480, 480), (720, 1280), (1080, 810), (500, 375), (640, 640)]
    out = fixed[:batch]
    while len(out) < batch:
        out.append((int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))))
    return out


# ---------------------------------------------------------------------------- evaluator inputs
def _name_seed(name: str) -> int:
    import hashlib
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little")


# (name, n_images, canvas, label count range, max dets)
EVAL_CASES = [
    ("e_small", 4, (640, 640), (0, 12), 40),
    ("e_dense", 3, (640, 640), (30, 60), 300),
    ("e_rect", 3, (512, 640), (1, 20), 100),
]


def eval_inputs(name):
    """Seeded detections / targets for the evaluator cases (shared with the tests):
    detections are jittered copies of the labels (so IoUs cover 0.3-1.0 and several detections
    compete for one label) plus random boxes; classes are drawn from 5 values."""
    _, n_img, canvas, (lo, hi), max_det = next(c for c in EVAL_CASES if c[0] == name)
    rng = np.random.Generator(np.random.PCG64(_name_seed(name)))
    H, W = canvas
    preds, targets, shapes = [], [], []
    for i in range(n_img):
        h0, w0 = int(rng.integers(200, 1400)), int(rng.integers(200, 1400))
        shapes.append((h0, w0))
        m = int(rng.integers(lo, hi + 1))
        cxy = rng.random((m, 2), dtype=np.float32) * np.float32([W, H])
        wh = rng.random((m, 2), dtype=np.float32) * np.float32(150) + np.float32(20)
        cls = rng.integers(0, 5, m).astype(np.float32)
        lab = np.concatenate([cls[:, None], cxy / np.float32([W, H]), wh / np.float32([W, H])], 1).astype(np.float32)
        targets.append(np.concatenate([np.full((m, 1), i, np.float32), lab], 1))
        k = 0 if (i == 1 and lo == 0) else int(rng.integers(max_det // 2, max_det + 1))
        rows = []
        for j in range(k):
            if m and rng.random() < 0.7:
                l = int(rng.integers(0, m))
                jit = (rng.random(4, dtype=np.float32) - np.float32(0.5)) * np.float32(0.5) * np.concatenate([wh[l], wh[l]])
                x1, y1 = cxy[l] - wh[l] / 2 + jit[:2]
                x2, y2 = cxy[l] + wh[l] / 2 + jit[2:]
                c = cls[l] if rng.random() < 0.8 else np.float32(rng.integers(0, 5))
            else:
                x1, y1 = rng.random(2, dtype=np.float32) * np.float32([W - 60, H - 60])
                x2, y2 = x1 + rng.random(dtype=np.float32) * 200 + 5, y1 + rng.random(dtype=np.float32) * 200 + 5
                c = np.float32(rng.integers(0, 5))
            rows.append([x1, y1, x2, y2, rng.random(dtype=np.float32), c])
        p = np.asarray(rows, np.float32).reshape(-1, 6)
        p = p[np.argsort(-p[:, 4], kind="stable")]
        preds.append(p)
    return preds, np.concatenate(targets, 0).astype(np.float32), shapes, canvas


