"""f-3: the losses of utils/net_tools.py:478-623 on the targets produced by refine_groundtruth /
det_groundtruth, as CUDA kernels behind torch.autograd.Function (csrc/loss.cu).

    smooth_l1(x)                                   :478-489
    refine_loss(refine_out, refine_gt, refine_pos_mask)                         :492-516
    det_clf_loss(refine_out, clf_out, det_out, det_gt, det_pos_mask, det_labels, iou_all_layers)   :519-623

Values follow the reference within the float tolerance of the north star (1e-5 relative: TF's float32
reductions and exp / log kernels have their own rounding).

Gradients.  By default they flow to the head outputs only (refine_out for refine_loss, det_out and clf_out
for det_clf_loss) and targets, masks, labels and the IoU factor are constants.  The reference graph has no
stop_gradient between det_groundtruth and the losses, so TF back-propagates two more terms into refine_out:
  (1) det_loss through det_gt = (offset_gt - refine_out) * mask (utils/net_tools.py:471), a linear term
      equal to the det_out gradient on the masked anchors — reproduced by
      `det_clf_loss(..., reference_gradients=True)`;
  (2) clf_loss through iou_factor (decode -> jaccard -> moments / min / max / pow, :590-600) — NOT reproduced
      (dropped in both modes); the integer masks and labels carry no gradient in TF either."""
from __future__ import annotations

import torch

from .. import _abi, config

__all__ = ["smooth_l1", "refine_loss", "det_clf_loss"]

NEGATIVE_RATIO = 3.0        # utils/net_tools.py:578


def smooth_l1(x):
    """0.5 * ((|x| - 1) * min(|x|, 1) + |x|)  (utils/net_tools.py:478-489); element-wise helper."""
    absx = torch.abs(x)
    return 0.5 * ((absx - 1) * torch.clamp(absx, max=1) + absx)


class _Shapes:
    """Layer layout derived from a per-layer list [B, fh, fw, A, inner] (no anchors needed here)."""

    def __init__(self, ts, inner):
        self.shapes, self.offsets = [], [0]
        for t in ts:
            if t.dim() < 3 or t.shape[-1] != inner:
                raise ValueError("per-layer tensors must be [B, ..., %d], got %s" % (inner, tuple(t.shape)))
            self.shapes.append(tuple(t.shape[1:-1]))
            self.offsets.append(self.offsets[-1] + t[0].numel() // inner)
        self.n_layers, self.n = len(self.shapes), self.offsets[-1]
        if not 1 <= self.n_layers <= _abi.MAX_LAYERS:
            raise ValueError("unsupported number of layers: %d" % self.n_layers)
        self.layout = _abi.Layout()
        self.layout.n_layers, self.layout.n_total = self.n_layers, self.n
        for i, o in enumerate(self.offsets):
            self.layout.offset[i] = o

    def split(self, flat, tail):
        """flat [B, N, tail] -> per-layer views [B, fh, fw, A, tail]."""
        return [flat[:, self.offsets[l]:self.offsets[l + 1]].reshape((flat.shape[0],) + s + (tail,)) for l, s in enumerate(self.shapes)]


def _table_of(ts, inner):
    return ts.table if isinstance(ts, _abi.LayerList) else _Shapes(list(ts), inner)


def _ints(ts):
    if isinstance(ts, _abi.LayerList) or all(t.dtype == torch.int32 for t in ts):
        return ts
    return [t.to(torch.int32) for t in ts]


def _floats(ts):
    if isinstance(ts, _abi.LayerList) or all(t.dtype == torch.float32 for t in ts):
        return ts
    return [t.to(torch.float32) for t in ts]


def _split_like(flat, xs, tail):
    out, off = [], 0
    for x in xs:
        n = x[0].numel() // tail
        out.append(flat[:, off:off + n].reshape(x.shape))
        off += n
    return out


class _SmoothL1Sum(torch.autograd.Function):
    """sum(smooth_l1((y - x) * mask)) / bs over all layers; gradient w.r.t. the x layers.  Tensors after
    the first n_layers of `xs` are "twins" (reference_gradients: the refine_out layers behind y) that
    receive the same gradient as the x layer of the same index."""

    @staticmethod
    def forward(ctx, y, mask, n_layers, *xs):
        ctx.n_twins = len(xs) - n_layers
        xs = xs[:n_layers]
        xs_d = [x.detach() for x in xs]
        dev = xs_d[0].device
        table = _table_of(y if isinstance(y, _abi.LayerList) else xs_d, 4)
        a, bb = _abi.DLArgs(), [-1]
        need_grad = any(ctx.needs_input_grad[3:])
        with _abi.device_guard(dev):
            xa = _abi.layered_arg(_floats(xs_d), table, 4, torch.float32, a, bb)
            ya = _abi.layered_arg(_floats(y), table, 4, torch.float32, a, bb)
            ma = _abi.layered_arg(_ints(mask), table, 1, torch.int32, a, bb)
            B = bb[0]
            out = torch.empty(1, dtype=torch.float32, device=dev)
            grad = torch.empty((B, table.n, 4), dtype=torch.float32, device=dev) if need_grad else None
            nb = int(_abi.lib.rod_smooth_l1_workspace_bytes(table.layout, B))
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            _abi.check(_abi.lib.rod_smooth_l1_loss(table.layout, ya, xa, ma, B, out.data_ptr(), grad.data_ptr() if need_grad else None,
                                                   1.0 / B, ws.data_ptr(), nb, _abi.stream_ptr(dev)))
        ctx.grad, ctx.shapes = grad, [x.shape for x in xs]
        return out[0]

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return (None, None, None) + (None,) * (len(ctx.shapes) + ctx.n_twins)
        flat = ctx.grad * g
        outs, off = [], 0
        for s in ctx.shapes:
            n = 1
            for d in s[1:-1]:
                n *= d
            outs.append(flat[:, off:off + n].reshape(s))
            off += n
        return (None, None, None) + tuple(outs) + tuple(outs[:ctx.n_twins])


class _ClfLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, labels, mask, iou, n_layers, *logits):
        lg = [x.detach() for x in logits]
        dev = lg[0].device
        C = lg[0].shape[-1]
        table = _table_of(mask if isinstance(mask, _abi.LayerList) else lg, C)
        a, bb = _abi.DLArgs(), [-1]
        with _abi.device_guard(dev):
            la = _abi.layered_arg(_floats(lg), table, C, torch.float32, a, bb)
            lb = _abi.layered_arg(_ints(labels), table, 1, torch.int32, a, bb)
            ma = _abi.layered_arg(_ints(mask), table, 1, torch.int32, a, bb)
            io = _abi.layered_arg(_floats(iou), table, 1, torch.float32, a, bb)
            B = bb[0]
            out = torch.empty(6, dtype=torch.float32, device=dev)
            nb = int(_abi.lib.rod_clf_loss_workspace_bytes(table.layout, B))
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            _abi.check(_abi.lib.rod_clf_loss(table.layout, la, lb, ma, io, B, C, NEGATIVE_RATIO, out.data_ptr(), ws.data_ptr(), nb,
                                             _abi.stream_ptr(dev)))
        ctx.saved = (table, lg, labels, mask, iou, ws, B, C)
        ctx.shapes = [x.shape for x in logits]
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g_aux):
        table, lg, labels, mask, iou, ws, B, C = ctx.saved
        dev = lg[0].device
        a, bb = _abi.DLArgs(), [-1]
        grad = torch.empty((B, table.n, C), dtype=torch.float32, device=dev)
        with _abi.device_guard(dev):
            la = _abi.layered_arg(_floats(lg), table, C, torch.float32, a, bb)
            lb = _abi.layered_arg(_ints(labels), table, 1, torch.int32, a, bb)
            ma = _abi.layered_arg(_ints(mask), table, 1, torch.int32, a, bb)
            io = _abi.layered_arg(_floats(iou), table, 1, torch.float32, a, bb)
            _abi.check(_abi.lib.rod_clf_loss_grad(table.layout, la, lb, ma, io, B, C, 1.0, ws.data_ptr(), grad.data_ptr(),
                                                  _abi.stream_ptr(dev)))
        grad = grad * g
        outs, off = [], 0
        for s in ctx.shapes:
            n = 1
            for d in s[1:-1]:
                n *= d
            outs.append(grad[:, off:off + n].reshape(s))
            off += n
        return (None, None, None, None) + tuple(outs)


def refine_loss(refine_out, refine_groundtruth, refine_pos_mask, dtype=torch.float32):
    """sum over layers of sum(smooth_l1((gt - out) * mask)) / bs  (utils/net_tools.py:492-516) -> 0-d tensor."""
    return _SmoothL1Sum.apply(refine_groundtruth, refine_pos_mask, len(refine_out), *list(refine_out)).to(dtype)


def det_clf_loss(refine_out, clf_out, det_out, det_groundtruth, det_pos_mask, det_labels, iou_all_layers,
                 dtype=torch.float32, return_details=False, reference_gradients=False):
    """(det_loss, clf_loss) of utils/net_tools.py:519-623: smooth-L1 on the ODM offsets, and the classification
    loss with hard-negative mining (negatives = the 3 * n_pos + bs anchors with the lowest background
    probability), positives weighted by the normalised IoU to the 4th power, clf_loss = neg_loss / 2 + pos_loss.
    return_details adds a dict with pos_loss, neg_loss, max_hard_pred, n_pos, n_neg (the reference's summaries).
    reference_gradients=True also sends det_loss's gradient into refine_out, as TF does through
    det_gt = (offset_gt - refine_out) * mask (see the module docstring; the IoU-factor path stays dropped)."""
    twins = list(refine_out) if reference_gradients else []
    if twins and len(twins) != len(det_out):
        raise ValueError("refine_out and det_out do not have the same number of layers")
    det_loss = _SmoothL1Sum.apply(det_groundtruth, det_pos_mask, len(det_out), *(list(det_out) + twins)).to(dtype)
    clf_loss, aux = _ClfLoss.apply(det_labels, det_pos_mask, iou_all_layers, len(clf_out), *list(clf_out))
    clf_loss = clf_loss.to(dtype)
    if return_details:
        return det_loss, clf_loss, {"pos_loss": aux[1], "neg_loss": aux[2], "max_hard_pred": aux[3], "n_pos": aux[4], "n_neg": aux[5]}
    return det_loss, clf_loss
