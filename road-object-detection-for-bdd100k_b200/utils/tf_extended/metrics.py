"""Mirror of `utils/tf_extended/metrics.py` for the evaluation chain of evaluate.py:162-197:
streaming_tp_fp_arrays (:133-204) -> precision_recall (:100-130) -> average_precision_voc07 / voc12
(:210-258).  The reference builds TF-1 streaming metrics (local variables + update ops); here the
"local variables" are device buffers owned by a StreamingTpFp object and every call of
streaming_tp_fp_arrays performs one update.  The accumulation, the precision / recall scan and both
AP integrals are CUDA kernels behind the C ABI (csrc/metrics.cu); the score sort in between uses
torch.sort (stable, descending = tf.nn.top_k's order: equal scores keep the lower index first)."""
from __future__ import annotations

import numpy as np
import torch

from ... import _abi

__all__ = ["precision_recall", "streaming_tp_fp_arrays", "average_precision_voc12", "average_precision_voc07",
           "precision_recall_values", "reset_local_variables", "StreamingTpFp"]

RM_THRESHOLD = 1e-4                                   # metrics.py:166
VOC07_RECALL_LEVELS = np.arange(0., 1.1, 0.1)         # metrics.py:251 (the very same float64 values)

_LOCAL_VARIABLES = {}                                 # metric scope name -> StreamingTpFp


def reset_local_variables():
    """tf.local_variables_initializer() for the metric scopes: drop every accumulated array."""
    _LOCAL_VARIABLES.clear()


def _u8(t):
    t = t.contiguous()
    return t.view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8)


class StreamingTpFp:
    """v_num_gbboxes, v_num_detections, v_tp, v_fp, v_scores (metrics.py:173-178) for `rows` classes,
    kept on the device; `ids` holds a global detection id per entry so that arrays gathered from
    several ranks can be put back into single-process order (rodet_b200.dist.allgather_tp_fp)."""

    def __init__(self, rows, device, capacity=4096):
        self.rows, self.device = int(rows), torch.device(device)
        self.capacity, self.upper, self.seen = 0, 0, 0
        self.count = torch.zeros(self.rows, dtype=torch.int64, device=self.device)       # v_num_detections
        self.nobjects = torch.zeros(self.rows, dtype=torch.int64, device=self.device)    # v_num_gbboxes
        self.scores = self.tp = self.fp = self.ids = None
        self._reserve(capacity)

    def _reserve(self, extra):
        need = self.upper + int(extra)
        if need <= self.capacity:
            return
        cap = max(need, 2 * self.capacity, 4096)
        new = (torch.empty((self.rows, cap), dtype=torch.float32, device=self.device),
               torch.empty((self.rows, cap), dtype=torch.uint8, device=self.device),
               torch.empty((self.rows, cap), dtype=torch.uint8, device=self.device),
               torch.empty((self.rows, cap), dtype=torch.int64, device=self.device))
        if self.capacity:
            for dst, src in zip(new, (self.scores, self.tp, self.fp, self.ids)):
                dst[:, :self.capacity].copy_(src)
        self.scores, self.tp, self.fp, self.ids = new
        self.capacity = cap

    def update(self, num_gbboxes, tp, fp, scores, remove_zero_scores=True, id_base=None):
        """One update op: scores / tp / fp [rows, ...] (flattened per row in order), num_gbboxes [rows, ...]."""
        scores = scores.to(torch.float32).reshape(self.rows, -1).contiguous()            # math_ops.to_float, reshape [-1]
        n = scores.shape[1]
        tp8, fp8 = _u8(tp.reshape(self.rows, -1)), _u8(fp.reshape(self.rows, -1))
        ngb = num_gbboxes.to(torch.int64).reshape(self.rows, -1).contiguous()            # math_ops.to_int64
        if tp8.shape[1] != n or fp8.shape[1] != n:
            raise ValueError("scores, tp and fp must hold the same number of detections")
        self._reserve(n)
        base = self.seen if id_base is None else int(id_base)
        P = lambda t: t.data_ptr() if t is not None and t.numel() else None
        with _abi.device_guard(self.device):
            _abi.check(_abi.lib.rod_tpfp_append(
                P(scores), P(tp8), P(fp8), self.rows, n, P(ngb), ngb.shape[1], 1 if remove_zero_scores else 0,
                RM_THRESHOLD, base, self.scores.data_ptr(), self.tp.data_ptr(), self.fp.data_ptr(), self.ids.data_ptr(),
                self.capacity, self.count.data_ptr(), self.nobjects.data_ptr(), _abi.stream_ptr(self.device)))
        self.upper += n
        self.seen += n

    def value(self, row=0, with_ids=False):
        """(v_num_gbboxes, v_num_detections, v_tp, v_fp, v_scores) of one row; reads the counter back."""
        k = int(self.count[row].item())
        self.upper = max(int(self.count.max().item()), 0)                                # tighten the host-side bound
        out = (self.nobjects[row], self.count[row].to(torch.int32), self.tp[row, :k].view(torch.bool),
               self.fp[row, :k].view(torch.bool), self.scores[row, :k])
        return out + (self.ids[row, :k],) if with_ids else out


def streaming_tp_fp_arrays(num_gbboxes, tp, fp, scores, remove_zero_scores=True, metrics_collections=None,
                           updates_collections=None, name=None, id_base=None):
    """Accumulates the TP / FP / score arrays and the ground-truth count over batches
    (utils/tf_extended/metrics.py:133-204).  Dict inputs (class -> tensor) return dicts.  Every call is
    one update; returns (value, update_op) with both equal to the state after it."""
    scope = name or 'streaming_tp_fp'
    if isinstance(scores, dict) or isinstance(fp, dict):
        keys = list(num_gbboxes.keys())
        st = _LOCAL_VARIABLES.get(scope)
        if st is None:
            st = _LOCAL_VARIABLES[scope] = StreamingTpFp(len(keys), scores[keys[0]].device)
            st.keys = keys
        elif getattr(st, "keys", None) != keys:
            raise ValueError("metric scope %r was created for classes %s" % (scope, getattr(st, "keys", None)))
        stack = lambda d: torch.stack([d[c].reshape(-1) for c in keys])
        st.update(stack(num_gbboxes), stack(tp), stack(fp), stack(scores), remove_zero_scores, id_base)
        vals = {c: st.value(i) for i, c in enumerate(keys)}
        return vals, dict(vals)
    st = _LOCAL_VARIABLES.get(scope)
    if st is None:
        st = _LOCAL_VARIABLES[scope] = StreamingTpFp(1, scores.device)
    st.update(num_gbboxes.reshape(1, -1), tp.reshape(1, -1), fp.reshape(1, -1), scores.reshape(1, -1), remove_zero_scores,
              id_base)
    val = st.value(0)
    return val, val


def precision_recall(num_gbboxes, num_detections, tp, fp, scores, dtype=torch.float64, scope=None):
    """Precision and recall arrays after sorting the detections by descending score
    (utils/tf_extended/metrics.py:100-130)."""
    if isinstance(scores, dict):
        d_precision, d_recall = {}, {}
        for c in num_gbboxes.keys():
            d_precision[c], d_recall[c] = precision_recall(num_gbboxes[c], num_detections[c], tp[c], fp[c], scores[c], dtype)
        return d_precision, d_recall
    scores = scores.reshape(-1)
    dev = scores.device
    k = int(num_detections)
    if k > scores.numel():
        raise ValueError("num_detections=%d exceeds the %d scores (tf.nn.top_k)" % (k, scores.numel()))
    # tf.nn.top_k(scores, k, sorted=True) + gather of tp / fp (:117-123): the library's own bitonic sort on
    # (score desc, index asc) composite keys, gather fused into its last pass
    scores = scores.to(torch.float32).contiguous()
    tp_u, fp_u = _u8(tp.reshape(-1)).contiguous(), _u8(fp.reshape(-1)).contiguous()
    tp_s, fp_s = torch.empty(k, dtype=torch.uint8, device=dev), torch.empty(k, dtype=torch.uint8, device=dev)
    if k:
        nbs = int(_abi.lib.rod_sort_scores_workspace_bytes(scores.numel()))
        wss = torch.empty(nbs, dtype=torch.uint8, device=dev)
        with _abi.device_guard(dev):
            _abi.check(_abi.lib.rod_sort_scores_desc(scores.data_ptr(), scores.numel(), k, tp_u.data_ptr(), fp_u.data_ptr(),
                                                     tp_s.data_ptr(), fp_s.data_ptr(), None, None, wss.data_ptr(), nbs,
                                                     _abi.stream_ptr(dev)))
    ngb = torch.as_tensor(num_gbboxes, device=dev).to(torch.int64).reshape(1).contiguous()
    precision = torch.empty(k, dtype=torch.float64, device=dev)
    recall = torch.empty(k, dtype=torch.float64, device=dev)
    if k:
        nb = int(_abi.lib.rod_precision_recall_workspace_bytes(k))
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        with _abi.device_guard(dev):
            _abi.check(_abi.lib.rod_precision_recall(tp_s.data_ptr(), fp_s.data_ptr(), k, ngb.data_ptr(), precision.data_ptr(),
                                                     recall.data_ptr(), ws.data_ptr(), nb, _abi.stream_ptr(dev)))
    if dtype != torch.float64:
        precision, recall = precision.to(dtype), recall.to(dtype)
    return [precision, recall]


def _average_precision(precision, recall):
    precision = precision.to(torch.float64).reshape(-1).contiguous()     # tf.cast(..., tf.float64)
    recall = recall.to(torch.float64).reshape(-1).contiguous()
    if precision.numel() != recall.numel():
        raise ValueError("precision and recall must have the same length")
    dev = precision.device
    out = torch.empty(2, dtype=torch.float64, device=dev)
    lv = (_abi.ctypes.c_double * 11)(*[float(t) for t in VOC07_RECALL_LEVELS])
    n = precision.numel()
    with _abi.device_guard(dev):
        _abi.check(_abi.lib.rod_average_precision(precision.data_ptr() if n else None, recall.data_ptr() if n else None, n, lv,
                                                  out.data_ptr(), _abi.stream_ptr(dev)))
    return out


def average_precision_voc12(precision, recall, name=None):
    """Area under the interpolated precision / recall curve (Pascal 2012 / ILSVRC), metrics.py:210-232."""
    return _average_precision(precision, recall)[1]


def average_precision_voc07(precision, recall, name=None):
    """11-point interpolated average precision (Pascal 2007), metrics.py:235-258."""
    return _average_precision(precision, recall)[0]


def precision_recall_values(xvals, precision, recall, name=None):
    """Precision at the given recall values (metrics.py:261-282); not on evaluate.py's path, plain torch."""
    from .math import cummax
    z, o = precision.new_zeros(1), recall.new_ones(1)
    precision = cummax(torch.cat([z, precision, z]), reverse=True)
    recall = torch.cat([recall.new_zeros(1), recall, o])
    return [precision[recall <= x].min() for x in xvals]
