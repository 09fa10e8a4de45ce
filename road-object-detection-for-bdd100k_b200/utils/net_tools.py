"""Drop-in for the hot-path functions of the reference's `utils/net_tools.py`
(anchors :21-142, encode/decode/jaccard :147-267, refine_groundtruth :270-428,
det_groundtruth :431-475, select :658-736, detected_bboxes :739-758).

Same names, positional/keyword parameters and defaults; CUDA `torch.Tensor`s where the
reference takes `tf.Tensor`s; the same list-per-layer / dict-per-class return structures.
`scope=` arguments are accepted and ignored.  Every function runs hand-written sm_100a
kernels through the C ABI (`include/rodet_b200.h`); nothing falls back to the CPU.

Extensions (not in the reference): `refine_groundtruth` also accepts a batch
(`center_bboxes[B,Gmax,4]`, `labels[B,Gmax]`, `gt_counts[B]`), since the reference runs it per
image and lets `tf.train.batch` stack the results (train.py:109-124);
`decode_detected_bboxes` fuses the decode call site (evaluate.py:139-143) into
`detected_bboxes`; `return_match_index=True` exposes the internal argmax."""
from __future__ import annotations

import collections
import math

import numpy as np
import torch

from .. import _abi, config
from ..anchor_table import AnchorTable, layer_table_for, table_for

__all__ = [
    "init_anchor", "n_anchor_each_layer", "anchors_one_layer", "anchors_all_layer",
    "encode_locations_one_layer", "decode_locations_one_layer", "jaccard", "refine_groundtruth",
    "det_groundtruth", "target_gen", "target_buffers", "bboxes_select_one_layer", "bboxes_select_all_layers", "detected_bboxes",
    "decode_detected_bboxes", "decode_locations_cascade", "detect_workspace", "detect_fallback_flags", "softmax",
]


# --------------------------------------------------------------------------------- anchors
def init_anchor(n_layers):
    """Pixel (height, width) of every anchor shape per layer, float64
    (utils/net_tools.py:21-82).  Reads `config.img_size`, `config.normal_anchor_range` and
    `config.special_anchor_range` like the reference."""
    img_h, img_w = config.img_size
    lo, hi = config.normal_anchor_range
    step = (hi - lo) / (n_layers - 1)
    root3 = math.sqrt(3)
    table = collections.OrderedDict()
    band_lo, band_hi = lo, lo + step
    for layer in range(n_layers):
        if layer == 0:
            scales = list(config.special_anchor_range)
        else:
            scales = [band_lo, (2 * band_lo + band_hi) / 3, (band_lo + 2 * band_hi) / 3]
            band_lo, band_hi = band_hi, band_hi + step
        shapes = []
        for s in scales:                         # ratios 1:1, 1:3 (wide), 3:1 (tall)
            shapes.append([s * img_h, s * img_w])
            shapes.append([s * img_h / root3, s * img_w * root3])
            shapes.append([s * img_h * root3, s * img_w / root3])
        px = np.array(shapes)
        px[:, 0] = np.minimum(px[:, 0], img_h)   # clip to the image
        px[:, 1] = np.minimum(px[:, 1], img_w)
        table["layer_%d" % (layer + 1)] = px
    return table


def n_anchor_each_layer(backbone_name):
    """[6, 9, 9, 9, 9, 9] (utils/net_tools.py:84-94)."""
    assert backbone_name in list(config.extract_feat_name.keys())
    shapes = init_anchor(len(config.extract_feat_name[backbone_name]))
    return [v.shape[0] for v in shapes.values()]


def anchors_one_layer(img_shape, feat_shape, anchors_one_layer, dtype=np.float32):
    """Cell centres y[fh,fw,1], x[fh,fw,1] and normalised sizes h[A], w[A]; float64 math,
    then cast (utils/net_tools.py:98-122).  Host NumPy, as in the reference."""
    rows, cols = np.mgrid[0:feat_shape[0], 0:feat_shape[1]]
    yc = ((rows + 0.5) / feat_shape[0])[..., None]
    xc = ((cols + 0.5) / feat_shape[1])[..., None]
    hh = anchors_one_layer[:, 0] / img_shape[0]
    ww = anchors_one_layer[:, 1] / img_shape[1]
    return yc.astype(dtype), xc.astype(dtype), hh.astype(dtype), ww.astype(dtype)


def anchors_all_layer(img_shape, feats_shape, anchors_all_layer):
    """List over layers of [y, x, h, w] (utils/net_tools.py:125-142)."""
    out = []
    for key, px in anchors_all_layer.items():
        out.append(list(anchors_one_layer(img_shape=img_shape, feat_shape=feats_shape[key],
                                          anchors_one_layer=px)))
    return out


# --------------------------------------------------------------------------------- helpers
def _f32(t, name):
    t = _abi.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32" % name)
    return t


def _check_list(ts, table, name):
    if len(ts) != table.n_layers:
        raise ValueError("%s has %d layers, anchors have %d" % (name, len(ts), table.n_layers))
    return list(ts)


_THR_CACHE = {}


def _thresholds(vals, table, name):
    key = (tuple(vals), table.n_layers)
    arr = _THR_CACHE.get(key)
    if arr is None:
        if len(key[0]) < table.n_layers:
            raise ValueError("%s has %d entries for %d layers" % (name, len(key[0]), table.n_layers))
        arr = _THR_CACHE[key] = _abi.float_array(key[0][:table.n_layers])
    return arr


# --------------------------------------------------------------------------------- encode / decode / jaccard
def encode_locations_one_layer(anchors_one_layer, center_bbox):
    """Offsets of ONE box [y, x, h, w] w.r.t. every anchor of a layer -> [fh, fw, A, 4]
    (utils/net_tools.py:147-179)."""
    box = _f32(center_bbox, "center_bbox").reshape(-1)
    if box.numel() != 4:
        raise ValueError("center_bbox must have 4 elements")
    t = layer_table_for(anchors_one_layer, box.device)
    out = torch.empty((t.n, 4), dtype=torch.float32, device=box.device)
    box = box.contiguous()
    with torch.cuda.device(box.device):
        _abi.check(_abi.lib.rod_encode_one_box(t.center.data_ptr(), 0, t.n, box.data_ptr(), out.data_ptr(),
                                               _abi.stream_ptr(box.device)))
    fh, fw, a = t.shapes[0]
    return out.view(fh, fw, a, 4)


def decode_locations_one_layer(anchors_one_layer, offset_bboxes):
    """Centre boxes [y, x, h, w] from offsets of shape [B, ..., 4] for one layer
    (utils/net_tools.py:182-234); the result has the input's shape."""
    off = _f32(offset_bboxes, "offset_bboxes")
    t = layer_table_for(anchors_one_layer, off.device)
    if off.dim() < 2 or off.shape[-1] != 4 or off[0].numel() != t.n * 4:
        raise ValueError("offset_bboxes must be [B, ..., 4] with %d anchors per image" % t.n)
    return _decode(t, [off], None, to_corner=False).view(off.shape)


def _decode(table, refine_out, det_out, to_corner):
    dev = refine_out[0].device
    B = refine_out[0].shape[0]
    out = torch.empty((B, table.n, 4), dtype=torch.float32, device=dev)
    if B == 0:
        return out
    a = _abi.DLArgs()
    with _abi.device_guard(dev):
        _abi.check(_abi.lib.rod_dl_decode(table.layout, a.one(table.center), a.many(refine_out),
                                          a.many(det_out), 1 if to_corner else 0, a.one(out),
                                          _abi.stream_ptr(dev)))
    return out


def decode_locations_cascade(anchors_all_layer, refine_out, det_out, to_corner=True):
    """Opt-in extra with NO reference counterpart (the reference decodes `refine_out + det_out` once, evaluate.py:141):
    the RefineDet cascade of BASELINE.json's north star — `det_out` decoded against the refined anchors
    `decode(anchors, refine_out)`, i.e. decode_locations_one_layer applied twice with its own corner -> re-derived-centre
    anchor step in between.  Returns the per-layer list of [B,fh,fw,A,4] boxes (corner form by default), ready for
    `detected_bboxes(predictions, localisations)`.  Results differ from the reference's; never used by default."""
    ro = [_f32(t, "refine_out") for t in refine_out]
    do = [_f32(t, "det_out") for t in det_out]
    dev = ro[0].device
    table = table_for(anchors_all_layer, dev)
    _check_list(ro, table, "refine_out")
    _check_list(do, table, "det_out")
    B = ro[0].shape[0]
    out = torch.empty((B, table.n, 4), dtype=torch.float32, device=dev)
    if B:
        a, bb = _abi.DLArgs(), [-1]
        with _abi.device_guard(dev):
            r = _abi.layered_arg(ro, table, 4, torch.float32, a, bb)
            d = _abi.layered_arg(do, table, 4, torch.float32, a, bb)
            _abi.check(_abi.lib.rod_decode_cascade(table.layout, table.center.data_ptr(), r, d, B, 1 if to_corner else 0,
                                                   out.data_ptr(), _abi.stream_ptr(dev)))
    return _abi.LayerList(out, table, True, False)


def jaccard(anchors, corner_bbox):
    """IoU of corner boxes `anchors[..., 4]` with `corner_bbox` ([4] or the same shape)
    -> [...]; plain divide, no safe-divide (utils/net_tools.py:237-267)."""
    a = _f32(anchors, "anchors").contiguous()
    g = _f32(corner_bbox, "corner_bbox").contiguous()
    n = a.numel() // 4
    bc = 1 if g.numel() == 4 and n != 1 else 0
    if not bc and g.numel() != a.numel():
        raise ValueError("corner_bbox must have 4 elements or the shape of anchors")
    out = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _abi.check(_abi.lib.rod_jaccard(a.data_ptr(), g.data_ptr(), bc, out.data_ptr(), n,
                                        _abi.stream_ptr(a.device)))
    return out


# --------------------------------------------------------------------------------- a9 ARM
def refine_groundtruth(anchors_all_layer, center_bboxes, labels, method, scope="refine_encode",
                       gt_counts=None, return_match_index=False, thresholds=None, forced_match=False):
    """ARM matching + encoding (utils/net_tools.py:270-428).

    Per image (reference form): center_bboxes[G,4], labels[G] -> four lists over layers of
    gt[fh,fw,A,4], cbboxes[fh,fw,A,4], labels[fh,fw,A,1] (int32), pos_mask[fh,fw,A,1] (int32).
    Batched extension: center_bboxes[B,Gmax,4], labels[B,Gmax], gt_counts[B] (int32, >= 1) ->
    the same lists with a leading batch dimension.
    forced_match=True (opt-in extra, NO reference counterpart, JACCARD_BIGGER only): every GT box additionally claims
    the anchor it overlaps best, whatever the layer threshold (SSD / RefineDet bipartite step; rod_arm_forced_match)."""
    if method == config.refine_method.JACCARD_TOPK:
        raise ValueError('Not support now')                       # :424
    if method not in (config.refine_method.NEAREST_NEIGHBOR, config.refine_method.JACCARD_BIGGER):
        raise ValueError('Function parameter "method" wrong')      # :426
    cb = _f32(center_bboxes, "center_bboxes")
    lab = _abi.require_cuda(labels, "labels")
    batched = cb.dim() == 3
    if not batched:
        if cb.dim() != 2:
            raise ValueError("center_bboxes must be [G,4] or [B,Gmax,4]")
        cb, lab = cb.unsqueeze(0), lab.unsqueeze(0)
    if cb.shape[-1] != 4 or lab.shape != cb.shape[:2]:
        raise ValueError("center_bboxes [.., G, 4] and labels [.., G] do not agree")
    if cb.shape[1] < 1:
        raise ValueError("at least one ground-truth box per image is required "
                         "(the reference indexes center_bboxes[0], utils/net_tools.py:398)")
    if lab.dtype not in (torch.int64, torch.int32):
        raise ValueError("labels must be int64 or int32")
    dev = cb.device
    table = table_for(anchors_all_layer, dev)
    thr = _thresholds(config.refine_pos_jac_val_all_layers if thresholds is None else thresholds,
                      table, "refine_pos_jac_val_all_layers")
    cb, lab = cb.contiguous(), lab.contiguous()
    if gt_counts is not None:
        gt_counts = _abi.require_cuda(gt_counts, "gt_counts").to(torch.int32).contiguous()
    B, N = cb.shape[0], table.n
    gt = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    cbo = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    lbo = torch.empty((B, N), dtype=torch.int32, device=dev)
    pos = torch.empty((B, N), dtype=torch.int32, device=dev)
    idx = torch.empty((B, N), dtype=torch.int32, device=dev) if return_match_index else None
    a = _abi.DLArgs()
    with _abi.device_guard(dev):
        _abi.check(_abi.lib.rod_dl_arm_match_encode(
            table.layout, a.one(table.corner), a.one(table.center), thr, a.one(cb), a.one(lab),
            a.one(gt_counts), int(method.value), a.one(gt), a.one(cbo), a.one(lbo), a.one(pos), a.one(idx),
            _abi.stream_ptr(dev)))
        if forced_match:
            if method != config.refine_method.JACCARD_BIGGER:
                raise ValueError("forced_match extends the JACCARD_BIGGER branch only")
            nb = int(_abi.lib.rod_arm_forced_match_workspace_bytes(B, cb.shape[1]))
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            _abi.check(_abi.lib.rod_arm_forced_match(
                table.layout, table.corner.data_ptr(), table.center.data_ptr(), cb.data_ptr(), lab.data_ptr(),
                1 if lab.dtype == torch.int64 else 0, gt_counts.data_ptr() if gt_counts is not None else None, B, cb.shape[1],
                gt.data_ptr(), cbo.data_ptr(), lbo.data_ptr(), pos.data_ptr(), idx.data_ptr() if idx is not None else None,
                ws.data_ptr(), nb, _abi.stream_ptr(dev)))
    LL = _abi.LayerList
    res = (LL(gt, table, batched, False), LL(cbo, table, batched, False), LL(lbo, table, batched, True),
           LL(pos, table, batched, True))
    if return_match_index:
        return res + (LL(idx, table, batched, False),)
    return res


# --------------------------------------------------------------------------------- a10 ODM
def det_groundtruth(refine_out, offset_gt, cbboxes, refine_labels, refine_pos_mask, anchors,
                    scope="det_encode", thresholds=None):
    """ODM target generation (utils/net_tools.py:431-475): four lists over layers of
    det_gt[B,fh,fw,A,4], mask[B,fh,fw,A,1] (int32), det_labels[B,fh,fw,A,1] (int32),
    iou[B,fh,fw,A]."""
    first = refine_out.flat if isinstance(refine_out, _abi.LayerList) else refine_out[0]
    dev = _abi.require_cuda(first, "refine_out").device
    table = table_for(anchors, dev)
    thr = _thresholds(config.det_pos_jac_val_all_layers if thresholds is None else thresholds,
                      table, "det_pos_jac_val_all_layers")
    a = _abi.DLArgs()
    bb = [-1]
    f32, i32 = torch.float32, torch.int32

    def ints(ts):      # the reference casts labels / masks to int32 (utils/net_tools.py:469)
        if isinstance(ts, _abi.LayerList) or all(t.dtype == i32 for t in ts):
            return ts
        return [t.to(i32) for t in ts]
    with _abi.device_guard(dev):
        ro = _abi.layered_arg(refine_out, table, 4, f32, a, bb)
        og = _abi.layered_arg(offset_gt, table, 4, f32, a, bb)
        cb = _abi.layered_arg(cbboxes, table, 4, f32, a, bb)
        lb = _abi.layered_arg(ints(refine_labels), table, 1, i32, a, bb)
        pm = _abi.layered_arg(ints(refine_pos_mask), table, 1, i32, a, bb)
        B, N = bb[0], table.n
        det_gt = torch.empty((B, N, 4), dtype=f32, device=dev)
        mask = torch.empty((B, N), dtype=i32, device=dev)
        dlab = torch.empty((B, N), dtype=i32, device=dev)
        iou = torch.empty((B, N), dtype=f32, device=dev)
        if B:
            _abi.check(_abi.lib.rod_odm_target(
                table.layout, table.center.data_ptr(), thr, ro, og, cb, lb, pm, B, det_gt.data_ptr(),
                mask.data_ptr(), dlab.data_ptr(), iou.data_ptr(), _abi.stream_ptr(dev)))
    LL = _abi.LayerList
    return (LL(det_gt, table, True, False), LL(mask, table, True, True), LL(dlab, table, True, True),
            LL(iou, table, True, False))


# --------------------------------------------------------------------------------- a9 + a10 fused
def target_buffers(anchors_all_layer, batch, device, need_cbboxes=True, match_index=False, flat=False):
    """Extension: preallocated flat outputs for `target_gen(..., out=...)` (a training loop reuses them every
    step instead of allocating 68 bytes per anchor per call).  flat=True carves all of them out of ONE uint8
    buffer (key "_flat"), so that a host consumer reads the whole result back with a single copy.  The per-layer
    views over these buffers are built by the first `target_gen(..., out=...)` call and handed out again by the
    later ones (the same list objects: 0.1 ms of view construction per call saved)."""
    dev = torch.device(device)
    table = table_for(anchors_all_layer, dev)
    B, N, f32, i32 = int(batch), table.n, torch.float32, torch.int32
    if flat:
        n4 = 3 if need_cbboxes else 2
        n1 = 4 + (1 if need_cbboxes else 0) + (1 if match_index else 0)
        buf = torch.empty((B * N * (16 * n4 + 4 * n1),), dtype=torch.uint8, device=dev)
        cursor = [0]

        def new(shape, dt):
            nbytes = 4
            for d in shape:
                nbytes *= d
            v = buf[cursor[0]:cursor[0] + nbytes].view(dt).view(shape)
            cursor[0] += nbytes
            return v
        order = ["gt", "det_gt"] + (["cb"] if need_cbboxes else []) + ["pos", "mask", "dlab", "iou"] + \
                (["lab"] if need_cbboxes else []) + (["idx"] if match_index else [])
        d = {k: None for k in ("gt", "pos", "cb", "lab", "idx", "det_gt", "mask", "dlab", "iou")}
        for k in order:
            d[k] = new((B, N, 4) if k in ("gt", "det_gt", "cb") else (B, N), f32 if k in ("gt", "det_gt", "cb", "iou") else i32)
        d["_flat"] = buf
        d["_sched"] = torch.zeros(2, dtype=i32, device=dev)
        d["_lists"] = None
        return d
    new = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
    return {"gt": new((B, N, 4), f32), "pos": new((B, N), i32),
            "cb": new((B, N, 4), f32) if need_cbboxes else None, "lab": new((B, N), i32) if need_cbboxes else None,
            "idx": new((B, N), i32) if match_index else None,
            "det_gt": new((B, N, 4), f32), "mask": new((B, N), i32), "dlab": new((B, N), i32), "iou": new((B, N), f32),
            "_sched": torch.zeros(2, dtype=i32, device=dev),     # scheduler counters: zero once, every call leaves them zero
            "_lists": None}                                       # per-layer views of these buffers, built by the first call


def target_gen(anchors_all_layer, center_bboxes, labels, refine_out, gt_counts=None,
               method=None, arm_thresholds=None, det_thresholds=None, return_match_index=False,
               need_cbboxes=True, out=None):
    """Extension: the training call sequence train.py:109-113 -> :147-149 in ONE kernel —
    `refine_groundtruth(anchors, center_bboxes, labels, JACCARD_BIGGER)` followed by
    `det_groundtruth(refine_out, refine_gt, refine_cbboxes, refine_labels, refine_pos_mask, anchors)`.
    Returns the two result tuples `((refine_gt, refine_cbboxes, refine_labels, refine_pos_mask),
    (det_gt, det_pos_mask, det_labels, iou_all_layers))`, bit-identical to the two calls; the matched GT,
    encoding, label and mask never round-trip through HBM (84 instead of 124 bytes per anchor).
    need_cbboxes=False skips writing refine_cbboxes / refine_labels (only det_groundtruth consumes them):
    the first tuple then holds None in their place.  out=target_buffers(...) writes into preallocated tensors."""
    JB = config.refine_method.JACCARD_BIGGER
    if method is not None and method != JB:
        if method == config.refine_method.JACCARD_TOPK:
            raise ValueError('Not support now')                   # utils/net_tools.py:424
        raise ValueError("target_gen fuses the JACCARD_BIGGER branch only; call refine_groundtruth + det_groundtruth")
    cb = _f32(center_bboxes, "center_bboxes")
    lab = _abi.require_cuda(labels, "labels")
    if cb.dim() != 3 or cb.shape[-1] != 4 or lab.shape != cb.shape[:2] or cb.shape[1] < 1:
        raise ValueError("target_gen takes a batch: center_bboxes [B,Gmax,4] (Gmax >= 1), labels [B,Gmax]")
    if lab.dtype not in (torch.int64, torch.int32):
        raise ValueError("labels must be int64 or int32")
    dev = cb.device
    table = table_for(anchors_all_layer, dev)
    ta = _thresholds(config.refine_pos_jac_val_all_layers if arm_thresholds is None else arm_thresholds, table,
                     "refine_pos_jac_val_all_layers")
    to = _thresholds(config.det_pos_jac_val_all_layers if det_thresholds is None else det_thresholds, table,
                     "det_pos_jac_val_all_layers")
    cb, lab = cb.contiguous(), lab.contiguous()
    if gt_counts is not None:
        gt_counts = _abi.require_cuda(gt_counts, "gt_counts").to(torch.int32).contiguous()
    B, N = cb.shape[0], table.n
    f32, i32 = torch.float32, torch.int32
    if out is None:
        out = target_buffers(table, B, dev, need_cbboxes, return_match_index)
    else:
        for k, v in out.items():
            if v is not None and not k.startswith("_") and isinstance(v, torch.Tensor) and (v.device != dev or v.shape[0] != B or v.shape[1] != N or not v.is_contiguous()):
                raise ValueError("out[%r] does not match batch %d x %d anchors on %s" % (k, B, N, dev))
        need_cbboxes = out["cb"] is not None and out["lab"] is not None
        return_match_index = out.get("idx") is not None
    gt, pos, cbo, lbo, idx = out["gt"], out["pos"], out["cb"] if need_cbboxes else None, out["lab"] if need_cbboxes else None, out.get("idx")
    det_gt, mask, dlab, iou = out["det_gt"], out["mask"], out["dlab"], out["iou"]
    a, bb = _abi.DLArgs(), [B]
    with _abi.device_guard(dev):
        ro = _abi.layered_arg(refine_out, table, 4, f32, a, bb)
        if B:
            _abi.check(_abi.lib.rod_target_fused(
                table.layout, table.corner.data_ptr(), table.center.data_ptr(), ta, to, cb.data_ptr(), lab.data_ptr(),
                1 if lab.dtype == torch.int64 else 0, gt_counts.data_ptr() if gt_counts is not None else None, B,
                cb.shape[1], ro, gt.data_ptr(), cbo.data_ptr() if cbo is not None else None,
                lbo.data_ptr() if lbo is not None else None, pos.data_ptr(), idx.data_ptr() if idx is not None else None,
                det_gt.data_ptr(), mask.data_ptr(), dlab.data_ptr(), iou.data_ptr(), out["_sched"].data_ptr(),
                _abi.stream_ptr(dev)))
    cached = out.get("_lists")
    if cached is not None and cached[0] is table and cached[1] == (need_cbboxes, return_match_index):
        return cached[2], cached[3]              # preallocated buffers: the per-layer views were built by the first call
    LL = _abi.LayerList
    arm = (LL(gt, table, True, False), LL(cbo, table, True, False) if need_cbboxes else None,
           LL(lbo, table, True, True) if need_cbboxes else None, LL(pos, table, True, True))
    if return_match_index:
        arm = arm + (LL(idx, table, True, False),)
    det = (LL(det_gt, table, True, False), LL(mask, table, True, True), LL(dlab, table, True, True),
           LL(iou, table, True, False))
    if "_sched" in out and "_lists" in out:      # (only dicts made by target_buffers carry the cache slot)
        out["_lists"] = (table, (need_cbboxes, return_match_index), arm, det)
    return arm, det


# --------------------------------------------------------------------------------- a11 select
def bboxes_select_one_layer(predictions_layer, localizations_layer, select_threshold=None,
                            num_classes=21, ignore_class=0, scope=None):
    """Per class c != ignore_class: scores = p_c * (p_c >= thr), bboxes = loc * (p_c >= thr)
    -> dicts c -> [B, N_l], c -> [B, N_l, 4] (utils/net_tools.py:658-697)."""
    return _select([predictions_layer], [localizations_layer], select_threshold, num_classes, ignore_class)


def bboxes_select_all_layers(predictions_net, localizations_net, select_threshold=None,
                             num_classes=21, ignore_class=0, scope=None):
    """The same over all layers, concatenated on the anchor axis (utils/net_tools.py:700-736)."""
    return _select(list(predictions_net), list(localizations_net), select_threshold, num_classes, ignore_class)


def _layout_from_lists(preds, inner):
    counts = []
    for p in preds:
        if p.dim() < 3 or p[0].numel() % inner:
            raise ValueError("per-layer tensors must be [B, ..., %d]" % inner)
        counts.append(p[0].numel() // inner)
    lay = _abi.Layout()
    lay.n_layers = len(counts)
    off = 0
    for i, c in enumerate(counts):
        lay.offset[i] = off
        off += c
    lay.offset[len(counts)] = off
    lay.n_total = off
    return lay, off


def _select(preds, locs, select_threshold, num_classes, ignore_class):
    thr = 0.0 if select_threshold is None else float(select_threshold)
    preds = [_f32(p, "predictions") for p in preds]
    locs = [_f32(l, "localizations") for l in locs]
    if len(preds) != len(locs) or len(preds) > _abi.MAX_LAYERS:
        raise ValueError("predictions / localizations layer lists do not agree")
    C = preds[0].shape[-1]
    if num_classes > C:
        raise ValueError("num_classes=%d exceeds the prediction depth %d" % (num_classes, C))
    dev, B = preds[0].device, preds[0].shape[0]
    lay, N = _layout_from_lists(preds, C)
    for p, l in zip(preds, locs):
        if l.shape[-1] != 4 or l[0].numel() // 4 != p[0].numel() // C or l.shape[0] != B:
            raise ValueError("predictions and localizations do not describe the same anchors")
    scores = torch.empty((C, B, N), dtype=torch.float32, device=dev)
    boxes = torch.empty((C, B, N, 4), dtype=torch.float32, device=dev)
    if B:
        pl, k1 = _abi.layered(preds, C)
        ll, k2 = _abi.layered(locs, 4)
        with torch.cuda.device(dev):
            _abi.check(_abi.lib.rod_bboxes_select(lay, pl, ll, B, C, ignore_class, thr, scores.data_ptr(),
                                                  boxes.data_ptr(), _abi.stream_ptr(dev)))
        del k1, k2
    d_scores, d_bboxes = {}, {}
    for c in range(num_classes):
        if c != ignore_class:
            d_scores[c] = scores[c]
            d_bboxes[c] = boxes[c]
    return d_scores, d_bboxes


# --------------------------------------------------------------------------------- a15 post-process
def detect_workspace(anchors_or_n_anchors, batch, top_k, device, num_classes=None):
    """Extension: a reusable workspace for detected_bboxes / decode_detected_bboxes (`workspace=`), so that a
    loop does not allocate one per call, and whose fallback flags can be read with detect_fallback_flags()."""
    C = config.total_obj_n if num_classes is None else int(num_classes)
    lay = _abi.Layout()
    if not isinstance(anchors_or_n_anchors, int):
        lay = table_for(anchors_or_n_anchors, torch.device(device)).layout
    nbytes = int(_abi.lib.rod_detect_workspace_bytes(lay, int(batch), C, int(top_k)))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
    _clean_detect_workspace(ws, int(batch), C, int(top_k))
    ws._rod_key = (int(batch), C, int(top_k))
    return ws


_WS_CACHE = {}                                   # (device, stream, batch, C, top_k) -> workspace; a handful of entries


def _default_detect_workspace(dev, need, B, C, top_k):
    """The workspace of calls that pass none: one per (device, stream, geometry), reused (calls on one stream are
    ordered, so they may share it) — except under CUDA-graph capture, where the buffer must belong to the graph."""
    if torch.cuda.is_current_stream_capturing():
        ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        _clean_detect_workspace(ws, B, C, top_k)
        return ws
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, B, C, top_k)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < need:
        if len(_WS_CACHE) >= 8:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))
        ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        _clean_detect_workspace(ws, B, C, top_k)
        _WS_CACHE[key] = ws
    return ws


def _clean_detect_workspace(ws, batch, C, top_k):
    """Zeroes the bookkeeping head of a detect workspace once (the library keeps it zero between calls: the sampled
    histogram is cleared by the kernels that consumed it, not by a memset in front of every call)."""
    off = int(_abi.lib.rod_detect_flags_offset(_abi.Layout(), batch, C, top_k))
    n = int(_abi.lib.rod_detect_workspace_clean_bytes(_abi.Layout(), batch, C, top_k))
    ws[off:off + n].zero_()


def detect_fallback_flags(workspace):
    """int32 [C, B] view into a detect_workspace(): non-zero where the last call handed the (class, image)
    segment to the exact general kernels (list overflow, sampled cut too high, massive ties).  Diagnostics
    only — results never depend on it."""
    B, C, k = workspace._rod_key
    off = int(_abi.lib.rod_detect_flags_offset(_abi.Layout(), B, C, k))
    return workspace[off:off + 4 * B * C].view(torch.int32).view(C, B)


def _detect(preds, locs, refine_out, det_out, anchors, select_threshold, nms_threshold, clipping_bbox,
            top_k, keep_top_k, num_classes, return_counts, from_logits=False, workspace=None, counts_out=None):
    thr = 0.0 if select_threshold is None else float(select_threshold)
    preds = [_f32(p, "predictions") for p in preds]
    dev, B, C = preds[0].device, preds[0].shape[0], preds[0].shape[-1]
    if num_classes > C:
        raise ValueError("num_classes=%d exceeds the prediction depth %d" % (num_classes, C))
    if anchors is not None:
        table = table_for(anchors, dev)
        lay, N, center = table.layout, table.n, table.center
        _check_list(preds, table, "predictions")
    else:
        lay, N = _layout_from_lists(preds, C)
        center = None
    if top_k > N:
        raise ValueError("top_k=%d must be <= the number of anchors %d (tf.nn.top_k)" % (top_k, N))
    scores = torch.empty((C, B, keep_top_k), dtype=torch.float32, device=dev)
    boxes = torch.empty((C, B, keep_top_k, 4), dtype=torch.float32, device=dev)
    counts = torch.empty((C, B), dtype=torch.int32, device=dev) if return_counts else None
    if counts_out is not None:                   # caller-provided [C, B] int32 (e.g. a slice of an all-gather staging buffer)
        if counts_out.shape != (C, B) or counts_out.dtype != torch.int32 or not counts_out.is_contiguous() or counts_out.device != dev:
            raise ValueError("counts_out must be a contiguous int32 [%d, %d] tensor on %s" % (C, B, dev))
        counts, return_counts = counts_out, True
    if B:
        need = int(_abi.lib.rod_detect_workspace_bytes(lay, B, C, top_k))
        ws = workspace
        if ws is None:
            ws = _default_detect_workspace(dev, need, B, C, int(top_k))
        elif ws.dtype != torch.uint8 or ws.numel() < need or ws.device != dev or not ws.is_contiguous():
            raise ValueError("workspace must be a contiguous uint8 CUDA tensor of >= %d bytes (detect_workspace())" % need)
        clip = None
        if clipping_bbox is not None:
            clip = torch.as_tensor(clipping_bbox, dtype=torch.float32, device=dev).reshape(4).contiguous()
        a = _abi.DLArgs()
        with _abi.device_guard(dev):
            entry = _abi.lib.rod_dl_detect_logits if from_logits else _abi.lib.rod_dl_detect
            _abi.check(entry(
                lay, a.one(center), a.many(preds), a.many(locs), a.many(refine_out), a.many(det_out), 0, thr,
                float(nms_threshold), int(top_k), int(keep_top_k), a.one(clip), a.one(scores), a.one(boxes),
                a.one(counts), a.one(ws), _abi.stream_ptr(dev)))
    rscores = {c: scores[c] for c in range(1, num_classes)}
    rbboxes = {c: boxes[c] for c in range(1, num_classes)}
    if return_counts:
        return rscores, rbboxes, counts
    return rscores, rbboxes


def softmax(clf_out, out=None):
    """slim.softmax over the class axis for one tensor or a per-layer list (evaluate.py:136-137,
    predict.py:127-128).  `out` may alias the input.  Same arithmetic as the softmax fused into
    detected_bboxes(..., from_logits=True), so both routes give identical detections."""
    if isinstance(clf_out, (list, tuple)):
        outs = out if out is not None else [None] * len(clf_out)
        return [softmax(t, o) for t, o in zip(clf_out, outs)]
    x = _f32(clf_out, "clf_out").contiguous()
    y = torch.empty_like(x) if out is None else out
    if x.numel():
        a = _abi.DLArgs()
        with _abi.device_guard(x.device):
            _abi.check(_abi.lib.rod_dl_softmax(a.one(x), a.one(y), _abi.stream_ptr(x.device)))
    return y


def detected_bboxes(predictions, localisations, select_threshold=None, nms_threshold=0.5,
                    clipping_bbox=None, top_k=800, keep_top_k=200, return_counts=False, from_logits=False,
                    workspace=None, counts_out=None):
    """select -> top_k -> per-class NMS -> zero-pad (-> clip) in two kernels
    (utils/net_tools.py:739-758).  predictions: list of [B,fh,fw,A,11] post-softmax scores (or the
    class logits with from_logits=True: slim.softmax is then fused into the select pass);
    localisations: list of [B,fh,fw,A,4] corner boxes.  Returns dicts c -> [B,keep_top_k],
    c -> [B,keep_top_k,4] for c = 1..config.total_obj_n-1."""
    locs = [_f32(l, "localisations") for l in localisations]
    return _detect(list(predictions), locs, None, None, None, select_threshold, nms_threshold,
                   clipping_bbox, top_k, keep_top_k, config.total_obj_n, return_counts, from_logits, workspace, counts_out)


def decode_detected_bboxes(anchors_all_layer, refine_out, det_out, predictions, select_threshold=None,
                           nms_threshold=0.5, clipping_bbox=None, top_k=800, keep_top_k=200,
                           return_counts=False, from_logits=False, workspace=None, counts_out=None):
    """Extension: the inference call sequence of evaluate.py:139-151 in one call —
    c2c(decode(anchors, refine_out + det_out)) is evaluated only for the top_k candidates of
    each (image, class) instead of materialising [B,N,4] boxes first."""
    ro = [_f32(t, "refine_out") for t in refine_out]
    do = [_f32(t, "det_out") for t in det_out]
    return _detect(list(predictions), None, ro, do, anchors_all_layer, select_threshold, nms_threshold,
                   clipping_bbox, top_k, keep_top_k, config.total_obj_n, return_counts, from_logits, workspace, counts_out)


# --------------------------------------------------------------------------------- f-3 losses
from .losses import smooth_l1, refine_loss, det_clf_loss  # noqa: E402,F401  (utils/net_tools.py:478-623)
