"""Deterministic synthetic BDD100K-shaped inputs for tests and bench.

Host-side NumPy only; `seed = 1234 + global_image_index` per image so that any
rank / shard regenerates exactly the images it owns.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 1234
MAX_GT = 100


def gt_boxes(image_index: int, max_gt: int = MAX_GT):
    """(corner[G,4] f32 in [0,1], labels[G] int64), 1 <= G <= max_gt, h,w >= 1e-3."""
    rng = np.random.default_rng(BASE_SEED + int(image_index))
    g = int(rng.integers(1, max_gt + 1))
    out = np.zeros((0, 4), dtype=np.float32)
    while out.shape[0] < g:
        m = 2 * (g - out.shape[0]) + 8
        c = rng.uniform(0.05, 0.95, size=(m, 2))
        hw = np.exp(rng.uniform(np.log(0.02), np.log(0.6), size=(m, 2)))
        cr = np.concatenate([c - hw / 2, c + hw / 2], axis=1)
        cr = np.clip(cr, 0.0, 1.0).astype(np.float32)
        ok = ((cr[:, 2] - cr[:, 0]) >= 1e-3) & ((cr[:, 3] - cr[:, 1]) >= 1e-3)
        out = np.concatenate([out, cr[ok]], axis=0)
    out = out[:g]
    labels = rng.integers(1, 11, size=g).astype(np.int64)
    return out, labels


def gt_batch(first_image: int, batch: int, max_gt: int = MAX_GT):
    """Padded corner boxes [B,max_gt,4] f32, labels [B,max_gt] int64, counts [B] i32."""
    boxes = np.zeros((batch, max_gt, 4), dtype=np.float32)
    labels = np.zeros((batch, max_gt), dtype=np.int64)
    counts = np.zeros((batch,), dtype=np.int32)
    for b in range(batch):
        cr, lb = gt_boxes(first_image + b, max_gt)
        boxes[b, :len(cr)] = cr
        labels[b, :len(cr)] = lb
        counts[b] = len(cr)
    return boxes, labels, counts


def head_offsets(image_index: int, n_anchors: int, salt: int = 0, sigma_c=0.1, sigma_s=0.2):
    """ARM / ODM head output for one image: [N,4] f32 ~ N(0, [sc,sc,ss,ss])."""
    rng = np.random.default_rng([BASE_SEED + int(image_index), 7919 + salt])
    o = rng.standard_normal(size=(n_anchors, 4)).astype(np.float32)
    o[:, :2] *= np.float32(sigma_c)
    o[:, 2:] *= np.float32(sigma_s)
    return o


def class_logits(image_index: int, n_anchors: int, n_cols: int = 11):
    """Class logits [N,11] f32 behind class_probs: 3*N(0,1) with +4 on background."""
    rng = np.random.default_rng([BASE_SEED + int(image_index), 104729])
    z = (rng.standard_normal(size=(n_anchors, n_cols)) * 3.0).astype(np.float32)
    z[:, 0] += np.float32(4.0)
    return z


def class_probs(image_index: int, n_anchors: int, n_cols: int = 11):
    """Post-softmax scores [N,11] f32: logits 3*N(0,1) with +4 on background."""
    z = class_logits(image_index, n_anchors, n_cols)
    z -= z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def stress_probs(image_index: int, n_anchors: int, thr: float = 0.3, n_cols: int = 11):
    """NMS stress (BASELINE config 5): every (anchor, class) score ~ U(thr, 1)."""
    rng = np.random.default_rng([BASE_SEED + int(image_index), 1299709])
    return rng.uniform(thr, 1.0, size=(n_anchors, n_cols)).astype(np.float32)


def _layer_grid(shapes):
    """Per flat anchor (layer-major, (fy, fx, a) row-major): layer, cell-centre y, x in [0,1]."""
    lay, ys, xs = [], [], []
    for l, (fh, fw, a) in enumerate(shapes):
        fy, fx, _ = np.meshgrid(np.arange(fh), np.arange(fw), np.arange(a), indexing="ij")
        lay.append(np.full(fh * fw * a, l, dtype=np.int32))
        ys.append(((fy.reshape(-1) + 0.5) / fh).astype(np.float32))
        xs.append(((fx.reshape(-1) + 0.5) / fw).astype(np.float32))
    return np.concatenate(lay), np.concatenate(ys), np.concatenate(xs)


_GRID_CACHE = {}


def clustered_probs(image_index: int, shapes, mode: str = "quadrant", thr: float = 0.3, n_cols: int = 11):
    """Spatially CLUSTERED class scores [N,11] f32 (what a trained head produces, unlike the i.i.d. generators
    above): `shapes` is the per-layer (fh, fw, A) list of the anchor layout.

      quadrant : the candidates (score >= thr) of class c all lie in ONE image quadrant of ONE of the three
                 finest layers, where about half of the anchors score U(thr, 1); everything else is U(0, thr/2).
      bumps    : a Gaussian bump of class `label` around every synthetic GT box of the image
                 (gt_boxes(image_index)), peak U(0.7, 1), width tied to the box size, on every layer;
                 background noise U(0, thr/2).  Hundreds of candidates per class packed around a few centres."""
    key = tuple(tuple(int(v) for v in s) for s in shapes)
    if key not in _GRID_CACHE:
        _GRID_CACHE[key] = _layer_grid(key)
    lay, ys, xs = _GRID_CACHE[key]
    n = lay.shape[0]
    rng = np.random.default_rng([BASE_SEED + int(image_index), 15485863, 0 if mode == "quadrant" else 1])
    p = rng.uniform(0.0, thr / 2, size=(n, n_cols)).astype(np.float32)
    if mode == "quadrant":
        for c in range(1, n_cols):
            l = int(rng.integers(0, min(3, len(key))))
            qy, qx = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            inside = (lay == l) & ((ys >= 0.5) == bool(qy)) & ((xs >= 0.5) == bool(qx))
            hit = inside & (rng.random(n) < 0.5)
            p[hit, c] = rng.uniform(thr, 1.0, size=int(hit.sum())).astype(np.float32)
    elif mode == "bumps":
        boxes, labels = gt_boxes(image_index)
        for (ymin, xmin, ymax, xmax), c in zip(boxes, labels):
            cy, cx, h, w = (ymin + ymax) / 2, (xmin + xmax) / 2, ymax - ymin, xmax - xmin
            d2 = ((ys - cy) / (0.25 * h + 1e-3)) ** 2 + ((xs - cx) / (0.25 * w + 1e-3)) ** 2
            bump = (rng.uniform(0.7, 1.0) * np.exp(-0.5 * d2) * rng.uniform(0.85, 1.0, size=n)).astype(np.float32)
            if int(c) < n_cols:
                p[:, int(c)] = np.maximum(p[:, int(c)], bump)
    else:
        raise ValueError("mode must be 'quadrant' or 'bumps'")
    return p
