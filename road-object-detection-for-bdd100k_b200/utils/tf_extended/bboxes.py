"""Drop-in for the hot-path functions of `utils/tf_extended/bboxes.py`:
bboxes_sort(:60-100), bboxes_sort_all_classes(:27-57), bboxes_clip(:103-136),
bboxes_resize(:139-163), bboxes_nms(:166-189), bboxes_nms_batch(:192-232),
bboxes_jaccard(:452-479), bboxes_intersection(:482-508).
CUDA torch tensors in, CUDA kernels underneath (include/rodet_b200.h); dict inputs are
treated per class like the reference does."""
from __future__ import annotations

import torch

from ... import _abi

__all__ = ["bboxes_sort_all_classes", "bboxes_sort", "bboxes_clip", "bboxes_resize", "bboxes_nms",
           "bboxes_nms_batch", "bboxes_jaccard", "bboxes_intersection", "bboxes_matching",
           "bboxes_matching_batch", "bboxes_filter_overlap"]


def _f32(t, name):
    t = _abi.require_cuda(t, name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32" % name)
    return t.contiguous()


def _sort(scores, bboxes, top_k, want_idx):
    s = _f32(scores, "scores")
    b = _f32(bboxes, "bboxes")
    if s.dim() != 2 or b.shape != s.shape + (4,):
        raise ValueError("bboxes_sort expects scores [B,N] and bboxes [B,N,4]")
    B, N = s.shape
    if not 1 <= top_k <= N:
        raise ValueError("top_k=%d must be in [1, N=%d] (tf.nn.top_k requires k <= N)" % (top_k, N))
    os_ = torch.empty((B, top_k), dtype=torch.float32, device=s.device)
    ob = torch.empty((B, top_k, 4), dtype=torch.float32, device=s.device)
    oi = torch.empty((B, top_k), dtype=torch.int32, device=s.device) if want_idx else None
    with torch.cuda.device(s.device):
        _abi.check(_abi.lib.rod_bboxes_sort(s.data_ptr(), b.data_ptr(), B, N, int(top_k), os_.data_ptr(),
                                            ob.data_ptr(), oi.data_ptr() if want_idx else None,
                                            _abi.stream_ptr(s.device)))
    return os_, ob, oi


def bboxes_sort_all_classes(classes, scores, bboxes, top_k=400, scope=None):
    """Sort by decreasing score keeping top_k; classes are gathered alongside."""
    s, b, idx = _sort(scores, bboxes, top_k, True)
    return torch.gather(classes, 1, idx.long()), s, b


def bboxes_sort(scores, bboxes, top_k=400, scope=None):
    """tf.nn.top_k(scores, k, sorted=True) + per-image gather of the boxes; equal scores keep
    the lower index first.  Dicts are processed per class."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        d_scores, d_bboxes = {}, {}
        for c in scores.keys():
            d_scores[c], d_bboxes[c] = bboxes_sort(scores[c], bboxes[c], top_k=top_k)
        return d_scores, d_bboxes
    s, b, _ = _sort(scores, bboxes, top_k, False)
    return s, b


def _ref_op(fn, bbox_ref, bboxes, out_inner):
    b = _f32(bboxes, "bboxes")
    r = torch.as_tensor(bbox_ref, dtype=torch.float32, device=b.device).contiguous()
    n = b.numel() // 4
    if r.numel() == 4:
        bc = 1
    elif r.numel() == b.numel():
        bc = 0
    else:
        raise ValueError("bbox_ref must be a single box or match bboxes")
    out = torch.empty(b.shape if out_inner == 4 else b.shape[:-1], dtype=torch.float32, device=b.device)
    with torch.cuda.device(b.device):
        _abi.check(fn(r.data_ptr(), bc, b.data_ptr(), out.data_ptr(), n, _abi.stream_ptr(b.device)))
    return out


def bboxes_clip(bbox_ref, bboxes, scope=None):
    """Intersect boxes with a reference box; empty boxes collapse (ymin=min(ymin,ymax))."""
    if isinstance(bboxes, dict):
        return {c: bboxes_clip(bbox_ref, bboxes[c]) for c in bboxes.keys()}
    return _ref_op(_abi.lib.rod_bboxes_clip, bbox_ref, bboxes, 4)


def bboxes_resize(bbox_ref, bboxes, name=None):
    """Express boxes in the frame of `bbox_ref` (which maps to [0,0,1,1])."""
    if isinstance(bboxes, dict):
        return {c: bboxes_resize(bbox_ref, bboxes[c]) for c in bboxes.keys()}
    b = _f32(bboxes, "bboxes")
    r = torch.as_tensor(bbox_ref, dtype=torch.float32, device=b.device).reshape(4).contiguous()
    out = torch.empty_like(b)
    with torch.cuda.device(b.device):
        _abi.check(_abi.lib.rod_bboxes_resize(r.data_ptr(), b.data_ptr(), out.data_ptr(), b.numel() // 4,
                                              _abi.stream_ptr(b.device)))
    return out


def bboxes_jaccard(bbox_ref, bboxes, name=None):
    """IoU of `bbox_ref` ((4,) or (N,4)) with bboxes (N,4) -> (N,), 0 where union <= 0."""
    return _ref_op(_abi.lib.rod_bboxes_jaccard, bbox_ref, bboxes, 1)


def bboxes_intersection(bbox_ref, bboxes, name=None):
    """Intersection area over box area -> (N,), 0 where the box area <= 0."""
    return _ref_op(_abi.lib.rod_bboxes_intersection, bbox_ref, bboxes, 1)


def _nms(scores, bboxes, nms_threshold, keep_top_k):
    s = _f32(scores, "scores")
    b = _f32(bboxes, "bboxes")
    if s.dim() != 2 or b.shape != s.shape + (4,):
        raise ValueError("bboxes_nms_batch expects scores [B,N] and bboxes [B,N,4]")
    B, N = s.shape
    os_ = torch.empty((B, keep_top_k), dtype=torch.float32, device=s.device)
    ob = torch.empty((B, keep_top_k, 4), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _abi.check(_abi.lib.rod_bboxes_nms_batch(s.data_ptr(), b.data_ptr(), B, N, float(nms_threshold),
                                                 int(keep_top_k), os_.data_ptr(), ob.data_ptr(), None,
                                                 _abi.stream_ptr(s.device)))
    return os_, ob


def bboxes_nms(scores, bboxes, nms_threshold=0.5, keep_top_k=200, scope=None):
    """Greedy NMS of one set: scores [N], bboxes [N,4] -> [keep_top_k], [keep_top_k,4],
    in selection order, zero padded."""
    s, b = _nms(scores.unsqueeze(0), bboxes.unsqueeze(0), nms_threshold, keep_top_k)
    return s[0], b[0]


def bboxes_nms_batch(scores, bboxes, nms_threshold=0.5, keep_top_k=200, scope=None):
    """Batched / per-class-dict NMS: [B,N], [B,N,4] -> [B,keep_top_k], [B,keep_top_k,4]."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        d_scores, d_bboxes = {}, {}
        for c in scores.keys():
            d_scores[c], d_bboxes[c] = bboxes_nms_batch(scores[c], bboxes[c], nms_threshold=nms_threshold,
                                                        keep_top_k=keep_top_k)
        return d_scores, d_bboxes
    return _nms(scores, bboxes, nms_threshold, keep_top_k)


def _matching(label, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold):
    s = _f32(scores, "scores")
    b = _f32(bboxes, "bboxes")
    gb = _f32(gbboxes, "gbboxes")
    gl = _abi.require_cuda(glabels, "glabels").contiguous()
    gd = _abi.require_cuda(gdifficults, "gdifficults").to(gl.dtype).contiguous()
    if gl.dtype not in (torch.int64, torch.int32):
        raise ValueError("glabels must be int64 or int32")
    if s.dim() != 2 or b.shape != s.shape + (4,) or gl.dim() != 2 or gb.shape != gl.shape + (4,) or \
            gl.shape[0] != s.shape[0] or gd.shape != gl.shape:
        raise ValueError("bboxes_matching_batch expects scores [B,N], bboxes [B,N,4], glabels / gdifficults [B,G], gbboxes [B,G,4]")
    B, N = s.shape
    G = gl.shape[1]
    n_gb = torch.empty((B,), dtype=torch.int64, device=s.device)
    tp = torch.empty((B, N), dtype=torch.bool, device=s.device)
    fp = torch.empty((B, N), dtype=torch.bool, device=s.device)
    with _abi.device_guard(s.device):
        _abi.check(_abi.lib.rod_bboxes_matching_batch(int(label), s.data_ptr(), b.data_ptr(), gl.data_ptr(), gb.data_ptr(),
                                                      gd.data_ptr(), 1 if gl.dtype == torch.int64 else 0, B, N, G,
                                                      float(matching_threshold), n_gb.data_ptr(), tp.data_ptr(),
                                                      fp.data_ptr(), _abi.stream_ptr(s.device)))
    return n_gb, tp, fp


def bboxes_matching(label, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5, scope=None):
    """TP / FP matching of one image's detections of class `label` (in score order) against its
    ground truth (utils/tf_extended/bboxes.py:246-334).  Returns (n_gbboxes, tp[N], fp[N])."""
    n, tp, fp = _matching(label, scores.unsqueeze(0), bboxes.unsqueeze(0), glabels.unsqueeze(0),
                          gbboxes.unsqueeze(0), gdifficults.unsqueeze(0), matching_threshold)
    return n[0], tp[0], fp[0]


def bboxes_matching_batch(labels, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5,
                          scope=None):
    """Batched / per-class-dict form (utils/tf_extended/bboxes.py:337-380); dict inputs return
    (d_n_gbboxes, d_tp, d_fp, scores) like the reference."""
    if isinstance(scores, dict) or isinstance(bboxes, dict):
        d_n_gbboxes, d_tp, d_fp = {}, {}, {}
        for c in labels:
            n, tp, fp, _ = bboxes_matching_batch(c, scores[c], bboxes[c], glabels, gbboxes, gdifficults,
                                                 matching_threshold)
            d_n_gbboxes[c], d_tp[c], d_fp[c] = n, tp, fp
        return d_n_gbboxes, d_tp, d_fp, scores
    n, tp, fp = _matching(labels, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold)
    return n, tp, fp, scores


def _gt_update(labels, bboxes, counts, distort_bbox, mirror, filter_overlap, threshold, assign_negative, clamp01):
    b = _f32(bboxes, "bboxes").contiguous()
    if b.dim() != 3 or b.shape[-1] != 4:
        raise ValueError("bboxes must be [B,G,4]")
    B, G = b.shape[0], b.shape[1]
    lab = labels.contiguous()
    if lab.dtype not in (torch.int64, torch.int32) or tuple(lab.shape) != (B, G) or lab.device != b.device:
        raise ValueError("labels must be an int64 / int32 [B,G] tensor on the boxes' device")
    dev = b.device
    cnt = None if counts is None else counts.to(device=dev, dtype=torch.int32).contiguous()
    crop = None if distort_bbox is None else torch.as_tensor(distort_bbox, dtype=torch.float32, device=dev).reshape(B, 4).contiguous()
    mir = None if mirror is None else torch.as_tensor(mirror, device=dev).reshape(B).to(torch.uint8).contiguous()
    ob, ol = torch.empty_like(b), torch.empty_like(lab)
    oc = torch.empty(B, dtype=torch.int32, device=dev)
    P = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(dev):
        _abi.check(_abi.lib.rod_gt_boxes_update(b.data_ptr(), lab.data_ptr(), 1 if lab.dtype == torch.int64 else 0, P(cnt), B, G,
                                                P(crop), P(mir), 1 if filter_overlap else 0, float(threshold),
                                                1 if assign_negative else 0, 1 if clamp01 else 0, ob.data_ptr(), ol.data_ptr(),
                                                oc.data_ptr(), _abi.stream_ptr(dev)))
    return ol, ob, oc


def bboxes_filter_overlap(labels, bboxes, threshold=0.5, assign_negative=False, scope=None):
    """Drops (or, with assign_negative, negates the label of) the boxes whose overlap with [0,0,1,1],
    relative to their own area, is not above `threshold` (utils/tf_extended/bboxes.py:408-428).
    labels [G], bboxes [G,4] -> filtered labels, bboxes (the kept count is read back, like
    tf.boolean_mask's dynamic shape)."""
    ol, ob, oc = _gt_update(labels.unsqueeze(0), bboxes.unsqueeze(0), None, None, None, True, threshold, assign_negative, False)
    if assign_negative:
        return ol[0], ob[0]
    k = int(oc[0].item())
    return ol[0, :k], ob[0, :k]
