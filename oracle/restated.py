"""Tier B oracle: vectorised NumPy restatement of the reference's box-level hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Never imported by the product.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Arithmetic is float32, one rounding per reference op, in the
reference's op order (TF's Eigen CPU kernels never contract a*b+c into an FMA).

Pinning status: checked bit-for-bit (masks, indices, labels, matched boxes, keep
sets) and to <=1 ulp (exp/log outputs) against Tier A = the UNMODIFIED reference
sources executed over `oracle/tf_shim` (tests/test_oracle_golden.py against the
committed fixtures in tests/golden/, and live in tests/test_oracle_vs_reference.py
when /root/reference is present).  The reference has no tests or golden vectors
of its own, and `tf.nn.top_k`, `tf.argmax`, `tf.image.non_max_suppression`,
`tf.exp`, `tf.log` live in TensorFlow (absent, version unpinned): for those five
the adopted semantics are the documented ones written out in oracle/tf_shim.

exp / log policy: the oracle evaluates them as float32(round(f64 libm)), i.e. the
correctly rounded float32 result (up to double rounding, p ~ 2^-29).  TF/Eigen and
NumPy's float32 kernels are within 1 ulp of that; the CUDA path computes the same
double-precision form so masks that depend on decoded boxes stay bit-stable.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

f32 = np.float32

# config.py:15-16,79-80,87
NORMAL_ANCHOR_RANGE = (0.05, 0.7)
SPECIAL_ANCHOR_RANGE = (0.02, 0.03)
REFINE_POS_JAC = (0.2, 0.3, 0.4, 0.4, 0.3, 0.3)
DET_POS_JAC = (0.5, 0.6, 0.7, 0.7, 0.6, 0.6)
TOTAL_OBJ_N = 11
FEAT_SIZES_418 = ((53, 53), (27, 27), (14, 14), (7, 7), (4, 4), (2, 2))   # config.py:34-35
FEAT_SIZES_512 = ((64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2))


# --------------------------------------------------------------------------- #
# a1-a4  anchors  (utils/net_tools.py:21-142)
# --------------------------------------------------------------------------- #
def init_anchor(n_layers, img_size, normal_range=NORMAL_ANCHOR_RANGE,
                special_range=SPECIAL_ANCHOR_RANGE):
    """utils/net_tools.py:21-82 — pixel [h, w] per anchor, float64."""
    boxes = OrderedDict()
    amin, amax = normal_range
    per = (amax - amin) / (n_layers - 1)
    rmin = amin
    rmax = rmin + per
    h, w = img_size
    s3 = math.sqrt(3)
    for i in range(n_layers):
        if i == 0:
            scales = [special_range[0], special_range[1]]
        else:
            scales = [rmin, (2 * rmin + rmax) / 3, (rmin + 2 * rmax) / 3]
            rmin = rmax
            rmax = rmin + per
        rows = []
        for s in scales:
            rows += [[s * h, s * w], [s * h / s3, s * w * s3], [s * h * s3, s * w / s3]]
        a = np.array(rows)
        a[:, 0] = np.minimum(a[:, 0], h)
        a[:, 1] = np.minimum(a[:, 1], w)
        boxes["layer_%d" % (i + 1)] = a
    return boxes


def anchors_one_layer(img_shape, feat_shape, sizes_px, dtype=np.float32):
    """utils/net_tools.py:98-122 — float64 math, then astype(float32)."""
    y, x = np.mgrid[0:feat_shape[0], 0:feat_shape[1]]
    xc = (x + 0.5) / feat_shape[1]
    yc = (y + 0.5) / feat_shape[0]
    hh = sizes_px[:, 0] / img_shape[0]
    ww = sizes_px[:, 1] / img_shape[1]
    return (np.expand_dims(yc, -1).astype(dtype), np.expand_dims(xc, -1).astype(dtype),
            hh.astype(dtype), ww.astype(dtype))


def anchors_all_layer(img_shape, feat_sizes, anchor_sizes=None):
    """utils/net_tools.py:125-142 — list over layers of [y, x, h, w]."""
    if anchor_sizes is None:
        anchor_sizes = init_anchor(len(feat_sizes), img_shape)
    out = []
    for i, (key, val) in enumerate(anchor_sizes.items()):
        fs = feat_sizes[key] if isinstance(feat_sizes, dict) else feat_sizes[i]
        out.append(list(anchors_one_layer(img_shape, fs, val)))
    return out


class AnchorTable:
    """Flat, layer-major then (fy, fx, a) row-major view of the anchors
    (flatten order of utils/net_tools.py:200,220-223,679-682,731-735)."""

    def __init__(self, anchors_all):
        corners, centers, shapes = [], [], []
        for y, x, h, w in anchors_all:
            y, x, h, w = (np.asarray(v, dtype=f32) for v in (y, x, h, w))
            # corner form, utils/net_tools.py:156-165 / 203-211 / 385-394 (f32)
            ymin = y - h / f32(2.)
            xmin = x - w / f32(2.)
            ymax = y + h / f32(2.)
            xmax = x + w / f32(2.)
            # re-derived centre form, utils/net_tools.py:168-171 / 214-217
            acy = (ymax + ymin) / f32(2.)
            acx = (xmax + xmin) / f32(2.)
            ah = ymax - ymin
            aw = xmax - xmin
            corners.append(np.stack([ymin, xmin, ymax, xmax], -1).reshape(-1, 4))
            centers.append(np.stack([acy, acx, ah, aw], -1).reshape(-1, 4))
            shapes.append((y.shape[0], y.shape[1], h.shape[0]))
        self.shapes = shapes                               # (fh, fw, A) per layer
        self.counts = [s[0] * s[1] * s[2] for s in shapes]
        self.offsets = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)
        self.n = int(self.offsets[-1])
        self.corner = np.concatenate(corners).astype(f32)  # [N,4] ymin,xmin,ymax,xmax
        self.center = np.concatenate(centers).astype(f32)  # [N,4] acy,acx,ah,aw
        self.layer_of = np.repeat(np.arange(len(shapes)), self.counts).astype(np.int32)

    def split(self, flat, tail=()):
        """[..., N, *tail] -> list of [..., fh, fw, A, *tail] per layer."""
        out = []
        lead = flat.shape[:flat.ndim - 1 - len(tail)]
        for l, (fh, fw, a) in enumerate(self.shapes):
            sl = flat[..., self.offsets[l]:self.offsets[l + 1], :] if tail else \
                flat[..., self.offsets[l]:self.offsets[l + 1]]
            out.append(sl.reshape(lead + (fh, fw, a) + tuple(tail)))
        return out

    @staticmethod
    def flatten(per_layer, tail_dims):
        """list of [(B,) fh, fw, A, *tail] -> [(B,) N, *tail] (tail_dims = len(tail))."""
        outs = []
        for t in per_layer:
            t = np.asarray(t)
            lead = t.ndim - 3 - tail_dims
            outs.append(t.reshape(t.shape[:lead] + (-1,) + t.shape[t.ndim - tail_dims:]
                                  if tail_dims else t.shape[:lead] + (-1,)))
        return np.concatenate(outs, axis=outs[0].ndim - 1 - tail_dims)


# --------------------------------------------------------------------------- #
# a5  box format helpers  (utils/common_tools.py:16-57)
# --------------------------------------------------------------------------- #
def center_to_corner(cb):
    cb = np.asarray(cb, dtype=f32)
    cy, cx, h, w = cb[..., 0], cb[..., 1], cb[..., 2], cb[..., 3]
    return np.stack([cy - h / f32(2), cx - w / f32(2), cy + h / f32(2), cx + w / f32(2)], -1)


def corner_to_center(cr):
    cr = np.asarray(cr, dtype=f32)
    ymin, xmin, ymax, xmax = cr[..., 0], cr[..., 1], cr[..., 2], cr[..., 3]
    return np.stack([(ymin + ymax) / f32(2.), (xmin + xmax) / f32(2.), ymax - ymin, xmax - xmin], -1)


def _exp32(x):
    with np.errstate(all="ignore"):
        return np.exp(np.asarray(x, dtype=np.float64)).astype(f32)


def _log32(x):
    with np.errstate(all="ignore"):
        return np.log(np.asarray(x, dtype=np.float64)).astype(f32)


# --------------------------------------------------------------------------- #
# a6-a8  encode / decode / jaccard  (utils/net_tools.py:147-267)
# --------------------------------------------------------------------------- #
def encode(center_anchor, center_bbox):
    """utils/net_tools.py:173-178 with the re-derived anchor centre form."""
    a = np.asarray(center_anchor, dtype=f32)
    g = np.asarray(center_bbox, dtype=f32)
    with np.errstate(all="ignore"):
        t_cy = (g[..., 0] - a[..., 0]) / a[..., 2]
        t_cx = (g[..., 1] - a[..., 1]) / a[..., 3]
        t_h = _log32(g[..., 2] / a[..., 2])
        t_w = _log32(g[..., 3] / a[..., 3])
    return np.stack([t_cy, t_cx, t_h, t_w], -1).astype(f32)


def decode(center_anchor, offsets):
    """utils/net_tools.py:226-231: cy = o0*ah + acy (mul, then add), h = exp(o2)*ah."""
    a = np.asarray(center_anchor, dtype=f32)
    o = np.asarray(offsets, dtype=f32)
    with np.errstate(all="ignore"):
        cy = o[..., 0] * a[..., 2] + a[..., 0]
        cx = o[..., 1] * a[..., 3] + a[..., 1]
        h = _exp32(o[..., 2]) * a[..., 2]
        w = _exp32(o[..., 3]) * a[..., 3]
    return np.stack([cy, cx, h, w], -1).astype(f32)


def jaccard(anchors_corner, corner_bbox):
    """utils/net_tools.py:254-266 — plain divide; union = (vol_a - inter) + area_g."""
    a = np.asarray(anchors_corner, dtype=f32)
    g = np.asarray(corner_bbox, dtype=f32)
    with np.errstate(all="ignore"):
        vol_a = (a[..., 3] - a[..., 1]) * (a[..., 2] - a[..., 0])
        iymin = np.maximum(a[..., 0], g[..., 0])
        ixmin = np.maximum(a[..., 1], g[..., 1])
        iymax = np.minimum(a[..., 2], g[..., 2])
        ixmax = np.minimum(a[..., 3], g[..., 3])
        h = np.maximum(iymax - iymin, f32(0.))
        w = np.maximum(ixmax - ixmin, f32(0.))
        inter = h * w
        union = vol_a - inter + (g[..., 2] - g[..., 0]) * (g[..., 3] - g[..., 1])
        return (inter / union).astype(f32)


# --------------------------------------------------------------------------- #
# a9  ARM matching + encode  (utils/net_tools.py:270-428)
# --------------------------------------------------------------------------- #
def arm_match_encode(table, center_bboxes, labels, thresholds=REFINE_POS_JAC,
                     method="JACCARD_BIGGER"):
    """One image.  Returns flat (gt[N,4] f32, cbboxes[N,4] f32, labels[N] i32,
    pos[N] i32, idx[N] i32).  JACCARD_BIGGER: utils/net_tools.py:382-421,316-343;
    NEAREST_NEIGHBOR: :354-380,283-312."""
    cb = np.asarray(center_bboxes, dtype=f32).reshape(-1, 4)
    lab = np.asarray(labels).astype(np.int32)          # :342 cast int64 -> int32
    g = cb.shape[0]
    assert g >= 1, "reference crashes on zero GT boxes (utils/net_tools.py:398)"
    n = table.n
    if method == "JACCARD_BIGGER":
        gcorner = center_to_corner(cb)                   # :323,398 (round trip)
        jac = jaccard(table.corner[None, :, :], gcorner[:, None, :])    # [G,N]
        best = jac.max(axis=0)                           # :405
        idx = jac.argmax(axis=0).astype(np.int32)        # :408 first max
        thr = np.asarray(thresholds, dtype=f32)[table.layer_of]
        pos = (best >= thr)                              # :406
    elif method == "NEAREST_NEIGHBOR":
        enc_all = encode(table.center[None, :, :], cb[:, None, :])      # [G,N,4]
        with np.errstate(all="ignore"):
            sq = enc_all * enc_all
            # tf.reduce_sum over the last axis of 4: Eigen sums sequentially
            dist = ((sq[..., 0] + sq[..., 1]) + sq[..., 2]) + sq[..., 3]
        idx = dist.argmin(axis=0).astype(np.int32)       # :365 first min
        pos = np.ones(n, dtype=bool)                     # :376
    elif method == "JACCARD_TOPK":
        raise ValueError("Not support now")              # :424
    else:
        raise ValueError('Function parameter "method" wrong')
    mg = cb[idx]                                         # matched GT centre box
    enc = encode(table.center, mg)
    posf = pos[:, None]
    gt = np.where(posf, enc, f32(0)).astype(f32)
    cbo = np.where(posf, mg, f32(0)).astype(f32)
    labo = np.where(pos, lab[idx], 0).astype(np.int32)
    return gt, cbo, labo, pos.astype(np.int32), idx


def forced_match(table, center_bboxes, labels, arm_out):
    """Opt-in extra without a reference counterpart: on top of arm_match_encode's outputs (one image) every GT box claims
    the anchor it overlaps best (argmax over anchors, first = lowest index, IoU > 0); several GT boxes on one anchor:
    highest IoU wins, ties lowest GT index.  Returns new (gt, cbboxes, labels, pos, idx)."""
    cb = np.asarray(center_bboxes, dtype=f32).reshape(-1, 4)
    lab = np.asarray(labels).astype(np.int32)
    gt, cbo, labo, pos, idx = [np.array(a, copy=True) for a in arm_out]
    jac = jaccard(table.corner[None, :, :], center_to_corner(cb)[:, None, :])     # [G,N]
    jac = np.where(np.isnan(jac), f32(0), jac)
    best_n, best_v = jac.argmax(axis=1), jac.max(axis=1)
    for g in range(cb.shape[0]):
        if not best_v[g] > 0:
            continue
        rivals = [o for o in range(cb.shape[0]) if o != g and best_v[o] > 0 and best_n[o] == best_n[g] and
                  (best_v[o] > best_v[g] or (best_v[o] == best_v[g] and o < g))]
        if rivals:
            continue
        n = int(best_n[g])
        gt[n] = encode(table.center[n], cb[g]) + f32(0)
        cbo[n] = cb[g] + f32(0)
        labo[n], pos[n], idx[n] = lab[g], 1, g
    return gt, cbo, labo, pos, idx


# --------------------------------------------------------------------------- #
# a10  ODM target generation  (utils/net_tools.py:431-475)
# --------------------------------------------------------------------------- #
def odm_target(table, refine_out, offset_gt, cbboxes, refine_labels, refine_pos,
               thresholds=DET_POS_JAC):
    """Flat batched inputs [B,N,4]/[B,N].  Returns (det_gt[B,N,4] f32, mask[B,N]
    i32, det_labels[B,N] i32, iou[B,N] f32)."""
    ro = np.asarray(refine_out, dtype=f32)
    og = np.asarray(offset_gt, dtype=f32)
    cb = np.asarray(cbboxes, dtype=f32)
    lab = np.asarray(refine_labels, dtype=np.int32)
    pm = np.asarray(refine_pos, dtype=np.int32)
    ref_corner = center_to_corner(decode(table.center, ro))      # :459-460
    gt_corner = center_to_corner(cb)                              # :463
    iou = jaccard(ref_corner, gt_corner)                          # :465 elementwise
    thr = np.asarray(thresholds, dtype=f32)[table.layer_of]
    with np.errstate(all="ignore"):
        m = (iou >= thr).astype(np.int32) * pm                    # :468-469
        det_gt = (og - ro) * m.astype(f32)[..., None]             # :471
    return det_gt.astype(f32), m, lab * m, iou


# --------------------------------------------------------------------------- #
# a17  inference decode call site  (evaluate.py:139-143, predict.py:130-134)
# --------------------------------------------------------------------------- #
def decode_corner(table, refine_out, det_out):
    s = np.asarray(refine_out, dtype=f32) + np.asarray(det_out, dtype=f32)
    return center_to_corner(decode(table.center, s))


def decode_cascade_corner(table, refine_out, det_out):
    """Opt-in extra without a reference counterpart: det_out decoded against the refined anchors, i.e.
    decode_locations_one_layer applied twice (anchors -> corners -> re-derived centre form in between, :156-171)."""
    refined = corner_to_center(center_to_corner(decode(table.center, refine_out)))
    return center_to_corner(decode(refined, det_out))


# --------------------------------------------------------------------------- #
# a11  select  (utils/net_tools.py:658-736)
# --------------------------------------------------------------------------- #
def bboxes_select(probs, boxes, select_threshold=None, num_classes=TOTAL_OBJ_N,
                  ignore_class=0):
    """probs [B,N,C], boxes [B,N,4] -> dicts c -> [B,N], c -> [B,N,4]."""
    thr = 0.0 if select_threshold is None else select_threshold
    p = np.asarray(probs, dtype=f32)
    b = np.asarray(boxes, dtype=f32)
    d_s, d_b = {}, {}
    for c in range(num_classes):
        if c == ignore_class:
            continue
        sc = p[:, :, c]
        fmask = (sc >= f32(thr)).astype(f32)
        d_s[c] = sc * fmask
        d_b[c] = b * fmask[..., None]
    return d_s, d_b


# --------------------------------------------------------------------------- #
# a12  sort / top-k  (utils/tf_extended/bboxes.py:60-100)
# --------------------------------------------------------------------------- #
def topk_indices(scores, k):
    """tf.nn.top_k: descending, equal values -> lower index first."""
    s = np.asarray(scores, dtype=f32)
    assert k <= s.shape[-1], "top_k: k must be <= N"
    with np.errstate(all="ignore"):
        return np.argsort(-s, axis=-1, kind="stable")[..., :k].astype(np.int32)


def bboxes_sort(scores, bboxes, top_k=400):
    if isinstance(scores, dict):
        out = {c: bboxes_sort(scores[c], bboxes[c], top_k) for c in scores}
        return {c: v[0] for c, v in out.items()}, {c: v[1] for c, v in out.items()}
    idx = topk_indices(scores, top_k)
    s = np.take_along_axis(np.asarray(scores, dtype=f32), idx, axis=-1)
    b = np.take_along_axis(np.asarray(bboxes, dtype=f32), idx[..., None], axis=-2)
    return s, b


# --------------------------------------------------------------------------- #
# a13  NMS  (utils/tf_extended/bboxes.py:166-232, tensors.py:59-86; TF kernel
#      semantics per SURVEY.md §8c / oracle/tf_shim)
# --------------------------------------------------------------------------- #
def nms_iou_matrix(b):
    b = np.asarray(b, dtype=f32)
    ymin = np.minimum(b[:, 0], b[:, 2]); ymax = np.maximum(b[:, 0], b[:, 2])
    xmin = np.minimum(b[:, 1], b[:, 3]); xmax = np.maximum(b[:, 1], b[:, 3])
    with np.errstate(all="ignore"):
        area = (ymax - ymin) * (xmax - xmin)
        ih = np.maximum(np.minimum(ymax[:, None], ymax[None, :]) -
                        np.maximum(ymin[:, None], ymin[None, :]), f32(0))
        iw = np.maximum(np.minimum(xmax[:, None], xmax[None, :]) -
                        np.maximum(xmin[:, None], xmin[None, :]), f32(0))
        inter = ih * iw
        iou = inter / ((area[:, None] + area[None, :]) - inter)
    bad = (area[:, None] <= 0) | (area[None, :] <= 0)
    return np.where(bad, f32(0), iou).astype(f32)


def nms_indices(scores, bboxes, nms_threshold=0.5, keep_top_k=200):
    """Indices (selection order) kept by tf.image.non_max_suppression."""
    s = np.asarray(scores, dtype=f32)
    with np.errstate(all="ignore"):
        order = np.argsort(-s, kind="stable")
    b = np.asarray(bboxes, dtype=f32)[order]
    over = nms_iou_matrix(b) > f32(nms_threshold)
    n = len(order)
    dead = np.zeros(n, dtype=bool)
    sel = []
    for i in range(n):
        if len(sel) >= keep_top_k:
            break
        if dead[i]:
            continue
        sel.append(i)
        dead |= over[i]
    return order[np.asarray(sel, dtype=np.int64)].astype(np.int32)


def bboxes_nms(scores, bboxes, nms_threshold=0.5, keep_top_k=200):
    idx = nms_indices(scores, bboxes, nms_threshold, keep_top_k)
    s = np.zeros(max(keep_top_k, len(idx)), dtype=f32)
    b = np.zeros((max(keep_top_k, len(idx)), 4), dtype=f32)
    s[:len(idx)] = np.asarray(scores, dtype=f32)[idx]
    b[:len(idx)] = np.asarray(bboxes, dtype=f32)[idx]
    return s, b


def bboxes_nms_batch(scores, bboxes, nms_threshold=0.5, keep_top_k=200):
    if isinstance(scores, dict):
        out = {c: bboxes_nms_batch(scores[c], bboxes[c], nms_threshold, keep_top_k)
               for c in scores}
        return {c: v[0] for c, v in out.items()}, {c: v[1] for c, v in out.items()}
    rs, rb = [], []
    for i in range(len(scores)):
        s, b = bboxes_nms(scores[i], bboxes[i], nms_threshold, keep_top_k)
        rs.append(s); rb.append(b)
    return np.stack(rs), np.stack(rb)


# --------------------------------------------------------------------------- #
# a14, a16  clip / resize / jaccard / intersection  (utils/tf_extended/bboxes.py)
# --------------------------------------------------------------------------- #
def bboxes_clip(bbox_ref, bboxes):
    """utils/tf_extended/bboxes.py:103-136."""
    if isinstance(bboxes, dict):
        return {c: bboxes_clip(bbox_ref, v) for c, v in bboxes.items()}
    r = np.asarray(bbox_ref, dtype=f32)
    b = np.asarray(bboxes, dtype=f32)
    ymin = np.maximum(b[..., 0], r[..., 0]); xmin = np.maximum(b[..., 1], r[..., 1])
    ymax = np.minimum(b[..., 2], r[..., 2]); xmax = np.minimum(b[..., 3], r[..., 3])
    ymin = np.minimum(ymin, ymax); xmin = np.minimum(xmin, xmax)
    return np.stack([ymin, xmin, ymax, xmax], -1)


def bboxes_resize(bbox_ref, bboxes):
    """utils/tf_extended/bboxes.py:139-163."""
    if isinstance(bboxes, dict):
        return {c: bboxes_resize(bbox_ref, v) for c, v in bboxes.items()}
    r = np.asarray(bbox_ref, dtype=f32)
    b = np.asarray(bboxes, dtype=f32)
    v = np.stack([r[0], r[1], r[0], r[1]])
    s = np.stack([r[2] - r[0], r[3] - r[1], r[2] - r[0], r[3] - r[1]])
    with np.errstate(all="ignore"):
        return ((b - v) / s).astype(f32)


def _safe_divide(num, den):
    """utils/tf_extended/math.py:25-38."""
    with np.errstate(all="ignore"):
        return np.where(den > 0, num / den, np.zeros_like(num)).astype(f32)


def bboxes_jaccard(bbox_ref, bboxes):
    """utils/tf_extended/bboxes.py:452-479 — union = ((-inter) + area_b) + area_ref."""
    r = np.asarray(bbox_ref, dtype=f32)
    b = np.asarray(bboxes, dtype=f32)
    iymin = np.maximum(b[..., 0], r[..., 0]); ixmin = np.maximum(b[..., 1], r[..., 1])
    iymax = np.minimum(b[..., 2], r[..., 2]); ixmax = np.minimum(b[..., 3], r[..., 3])
    h = np.maximum(iymax - iymin, f32(0)); w = np.maximum(ixmax - ixmin, f32(0))
    inter = h * w
    union = (-inter + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
             + (r[..., 2] - r[..., 0]) * (r[..., 3] - r[..., 1]))
    return _safe_divide(inter, union)


def bboxes_intersection(bbox_ref, bboxes):
    """utils/tf_extended/bboxes.py:482-508."""
    r = np.asarray(bbox_ref, dtype=f32)
    b = np.asarray(bboxes, dtype=f32)
    iymin = np.maximum(b[..., 0], r[..., 0]); ixmin = np.maximum(b[..., 1], r[..., 1])
    iymax = np.minimum(b[..., 2], r[..., 2]); ixmax = np.minimum(b[..., 3], r[..., 3])
    h = np.maximum(iymax - iymin, f32(0)); w = np.maximum(ixmax - ixmin, f32(0))
    inter = h * w
    vol = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return _safe_divide(inter, vol)


# --------------------------------------------------------------------------- #
# a15  post-process driver  (utils/net_tools.py:739-758)
# --------------------------------------------------------------------------- #
def detected_bboxes(probs, boxes, select_threshold=None, nms_threshold=0.5,
                    clipping_bbox=None, top_k=800, keep_top_k=200,
                    num_classes=TOTAL_OBJ_N):
    """probs [B,N,C] (post-softmax), boxes [B,N,4] corner ->
    dicts c -> [B,keep], c -> [B,keep,4].

    Equivalent compacting evaluation is NOT used here: this
    is the literal select -> top_k -> NMS -> pad chain, so zero-score entries are
    real candidates exactly as in the reference."""
    d_s, d_b = bboxes_select(probs, boxes, select_threshold, num_classes)
    d_s, d_b = bboxes_sort(d_s, d_b, top_k=top_k)
    d_s, d_b = bboxes_nms_batch(d_s, d_b, nms_threshold, keep_top_k)
    if clipping_bbox is not None:
        d_b = bboxes_clip(clipping_bbox, d_b)
    return d_s, d_b


def softmax(logits):
    """slim.softmax (evaluate.py:136-137) — row max subtracted, float32."""
    a = np.asarray(logits, dtype=f32)
    e = np.exp(a - a.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(f32)


# --------------------------------------------------------------------------- #
# f-2  eval TP/FP matching  (utils/tf_extended/bboxes.py:246-380)
# --------------------------------------------------------------------------- #
def bboxes_matching(label, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5):
    """One image, one class: greedy matching of the detections (in the given order) to GT boxes.
    Returns (n_gbboxes int64, tp[N] bool, fp[N] bool).  utils/tf_extended/bboxes.py:267-334."""
    s = np.asarray(scores, dtype=f32).reshape(-1)
    b = np.asarray(bboxes, dtype=f32).reshape(-1, 4)
    gl = np.asarray(glabels)
    gb = np.asarray(gbboxes, dtype=f32).reshape(-1, 4)
    gd = np.asarray(gdifficults).astype(bool)
    same = (gl == gl.dtype.type(label))
    n_gb = np.int64(np.count_nonzero(same & ~gd))                        # :274-275
    gmatch = np.zeros(gl.shape, dtype=bool)
    tp = np.zeros(s.shape, dtype=bool)
    fp = np.zeros(s.shape, dtype=bool)
    for i in range(s.shape[0]):
        jac = bboxes_jaccard(b[i], gb) * same.astype(f32)               # :292-293
        idx = int(np.argmax(jac))                                        # :296 first max
        match = jac[idx] > f32(matching_threshold)                       # :298
        existing = gmatch[idx]
        not_diff = not gd[idx]
        tp[i] = not_diff and match and not existing                      # :304-305
        fp[i] = not_diff and (existing or not match)                     # :307-308
        if not_diff and match:                                           # :311-313
            gmatch[idx] = True
    return n_gb, tp, fp


def bboxes_matching_batch(labels, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5):
    """Batched / dict form (utils/tf_extended/bboxes.py:337-380): dict inputs return
    (d_n_gbboxes, d_tp, d_fp, scores)."""
    if isinstance(scores, dict):
        d_n, d_tp, d_fp = {}, {}, {}
        for c in labels:
            d_n[c], d_tp[c], d_fp[c], _ = bboxes_matching_batch(c, scores[c], bboxes[c], glabels, gbboxes,
                                                                  gdifficults, matching_threshold)
        return d_n, d_tp, d_fp, scores
    out = [bboxes_matching(labels, scores[i], bboxes[i], glabels[i], gbboxes[i], gdifficults[i], matching_threshold)
           for i in range(len(scores))]
    return (np.asarray([o[0] for o in out], dtype=np.int64), np.stack([o[1] for o in out]),
            np.stack([o[2] for o in out]), scores)


# --------------------------------------------------------------------------- #
# f-2  evaluation metrics  (utils/tf_extended/metrics.py:100-258, evaluate.py:162-197)
# --------------------------------------------------------------------------- #
class StreamingTpFp:
    """The local variables of tfe.streaming_tp_fp_arrays and its update op (metrics.py:157-195)."""

    def __init__(self):
        self.nobjects, self.ndetections = np.int64(0), np.int32(0)
        self.scores = np.zeros((0,), f32)
        self.tp = np.zeros((0,), bool)
        self.fp = np.zeros((0,), bool)

    def update(self, num_gbboxes, tp, fp, scores, remove_zero_scores=True):
        scores = np.asarray(scores, dtype=f32).reshape(-1)
        tp = np.asarray(tp).astype(bool).reshape(-1)
        fp = np.asarray(fp).astype(bool).reshape(-1)
        if remove_zero_scores:                                   # the mask is only applied in this branch (:164-170)
            mask = (tp | fp) & (scores > f32(1e-4))
            scores, tp, fp = scores[mask], tp[mask], fp[mask]
        self.nobjects = np.int64(self.nobjects + np.asarray(num_gbboxes, dtype=np.int64).sum())
        self.ndetections = np.int32(self.ndetections + scores.size)
        self.scores = np.concatenate([self.scores, scores])
        self.tp = np.concatenate([self.tp, tp])
        self.fp = np.concatenate([self.fp, fp])
        return self.value()

    def value(self):
        return self.nobjects, self.ndetections, self.tp, self.fp, self.scores


def precision_recall(num_gbboxes, num_detections, tp, fp, scores):
    """metrics.py:117-130: top_k sort (ties: lower index first), float64 cumsums, _safe_div."""
    scores = np.asarray(scores, dtype=f32).reshape(-1)
    k = int(num_detections)
    idx = np.argsort(-scores, kind="stable")[:k]
    tpc = np.cumsum(np.asarray(tp).reshape(-1)[idx].astype(np.float64))
    fpc = np.cumsum(np.asarray(fp).reshape(-1)[idx].astype(np.float64))
    ngb = np.float64(num_gbboxes)
    with np.errstate(all="ignore"):
        recall = np.where(ngb > 0, tpc / ngb, 0.0)
        precision = np.where(tpc + fpc > 0, tpc / (tpc + fpc), 0.0)
    return precision, recall


def cummax(x, reverse=False):
    """utils/tf_extended/math.py:41-67."""
    x = np.asarray(x)
    return np.maximum.accumulate(x[::-1])[::-1] if reverse else np.maximum.accumulate(x)


def average_precision_voc12(precision, recall):
    """metrics.py:210-232."""
    p = np.concatenate([[0.], np.asarray(precision, np.float64), [0.]])
    r = np.concatenate([[0.], np.asarray(recall, np.float64), [1.]])
    p = cummax(p, reverse=True)
    return np.float64(np.sum(p[1:] * (r[1:] - r[:-1])))


def average_precision_voc07(precision, recall):
    """metrics.py:235-258: 11 recall levels np.arange(0., 1.1, 0.1), each max / 11, summed in order."""
    p = np.concatenate([np.asarray(precision, np.float64), [0.]])
    r = np.concatenate([np.asarray(recall, np.float64), [np.inf]])
    ap = None
    for t in np.arange(0., 1.1, 0.1):
        v = np.max(p[r >= t]) / 11.
        ap = v if ap is None else ap + v
    return np.float64(ap)


# --------------------------------------------------------------------------- #
# f-4  ground-truth boxes of the training input pipeline
# (utils/data_pileline_tools.py:88-108, process.py:134-138, tf_image.py:284-289)
# --------------------------------------------------------------------------- #
def bboxes_filter_overlap(labels, bboxes, threshold=0.5, assign_negative=False):
    """utils/tf_extended/bboxes.py:408-428."""
    labels = np.asarray(labels)
    bboxes = np.asarray(bboxes, dtype=f32).reshape(-1, 4)
    scores = bboxes_intersection(np.asarray([0, 0, 1, 1], f32), bboxes)
    mask = scores > f32(threshold)
    if assign_negative:
        return np.where(mask, labels, -labels), bboxes
    return labels[mask], bboxes[mask]


def gt_boxes_train(labels, bboxes, distort_bbox=None, mirror=False, crop_overlap=0.3, assign_negative=False):
    """One image: resize to the crop, drop boxes mostly outside it, flip, clamp to [0,1]."""
    labels = np.asarray(labels)
    b = np.asarray(bboxes, dtype=f32).reshape(-1, 4)
    if distort_bbox is not None:
        b = bboxes_resize(np.asarray(distort_bbox, f32), b)
    if crop_overlap is not None:
        labels, b = bboxes_filter_overlap(labels, b, crop_overlap, assign_negative)
    if mirror:
        b = np.stack([b[:, 0], f32(1) - b[:, 3], b[:, 2], f32(1) - b[:, 1]], axis=-1)       # tf_image.py:286-288
    b = np.minimum(np.maximum(b, f32(0.)), f32(1.))                                        # data_pileline_tools.py:107-108
    return labels, b.astype(f32)


# --------------------------------------------------------------------------- #
# f-3  losses  (utils/net_tools.py:478-623)
# --------------------------------------------------------------------------- #
def smooth_l1(x):
    """:478-489."""
    a = np.abs(x)
    return 0.5 * ((a - 1) * np.minimum(a, 1) + a)


def smooth_l1_loss(y_layers, x_layers, mask_layers):
    """refine_loss (:492-516) / det_loss (:538-551): sum over layers of sum(smooth_l1((y - x) * mask)) / bs.
    float32 element-wise ops as the reference, float64 accumulation."""
    total = 0.0
    bs = np.asarray(x_layers[0]).shape[0]
    for y, x, m in zip(y_layers, x_layers, mask_layers):
        y, x = np.asarray(y, f32), np.asarray(x, f32)
        m = np.asarray(m).astype(f32).reshape(y.shape[:-1] + (1,))
        total += float(np.sum(smooth_l1((y - x) * m).astype(np.float64))) / bs
    return total


def clf_loss(clf_layers, det_pos_mask, det_labels, iou_layers, negative_ratio=3.0):
    """Classification half of det_clf_loss (:553-615).  Returns dict(clf_loss, pos_loss, neg_loss,
    max_hard_pred, n_pos, n_neg, weights) — weights [B*N...] in the reference's flatten order are the
    per-anchor cross-entropy weights (for the gradient check)."""
    bs = np.asarray(clf_layers[0]).shape[0]
    C = np.asarray(clf_layers[0]).shape[-1]
    logits = np.concatenate([np.asarray(t, f32).reshape(-1, C) for t in clf_layers], axis=0)
    gcls = np.concatenate([np.asarray(t).reshape(-1) for t in det_labels]).astype(np.int64)
    pmask = np.concatenate([np.asarray(t).reshape(-1) for t in det_pos_mask]).astype(bool)
    n_pos = int(pmask.sum())
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    s = e.sum(axis=1, keepdims=True)
    pred0 = (e[:, 0:1] / s)[:, 0].astype(f32)
    nmask = ~pmask
    nvalues = np.where(nmask, pred0, f32(1.0))
    max_neg = int(nmask.sum())
    n_neg = min(int(f32(negative_ratio) * f32(n_pos)) + bs, max_neg)
    max_hard = np.sort(nvalues, kind="stable")[n_neg - 1] if n_neg > 0 else f32(0)
    sel = nmask & (nvalues < max_hard)
    factors = []
    for iou in iou_layers:
        v = np.asarray(iou, f32)
        ax = tuple(range(1, v.ndim))
        mean = v.mean(axis=ax, keepdims=True, dtype=np.float64).astype(f32)
        var = ((v - mean) * (v - mean)).mean(axis=ax, keepdims=True, dtype=np.float64).astype(f32)
        v = (v - mean) / np.sqrt(var + f32(1e-8))
        v = v + (f32(0.) - v.min(axis=ax, keepdims=True))
        v = v / (v.max(axis=ax, keepdims=True) + f32(1e-8))
        factors.append((v ** 4).reshape(-1))
    factor = np.concatenate(factors).astype(f32)
    lse = np.log(s[:, 0].astype(np.float64))
    ce_lab = lse - z[np.arange(z.shape[0]), gcls].astype(np.float64)
    ce_0 = lse - z[:, 0].astype(np.float64)
    pos_loss = float(np.sum(ce_lab * pmask * factor.astype(np.float64))) / bs
    neg_loss = float(np.sum(ce_0 * sel)) / bs
    w = np.where(pmask, factor.astype(np.float64) / bs, np.where(sel, 0.5 / bs, 0.0))
    return dict(clf_loss=neg_loss / 2. + pos_loss, pos_loss=pos_loss, neg_loss=neg_loss, max_hard_pred=float(max_hard),
                n_pos=n_pos, n_neg=n_neg, weights=w, targets=np.where(pmask, gcls, 0), softmax=(e / s))
