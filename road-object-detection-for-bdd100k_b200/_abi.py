"""ctypes binding of include/rodet_b200.h plus the DLPack hand-off.

Fails loudly at import when `librodet_b200.so` is missing: this package has no CPU or
PyTorch fallback for any hot-path function."""
from __future__ import annotations

import ctypes
import os

import torch
from torch.utils.dlpack import to_dlpack

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librodet_b200.so")
MAX_LAYERS = 8
MAX_TOPK = 1024

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        "rodet_b200: %s not found. Build it with `make` (nvcc, sm_100a) or "
        "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
lib = ctypes.CDLL(LIB_PATH)


class Layout(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("n_total", ctypes.c_int32),
                ("offset", ctypes.c_int32 * (MAX_LAYERS + 1))]


class Layered(ctypes.Structure):
    _fields_ = [("base", ctypes.c_void_p * MAX_LAYERS), ("batch_stride", ctypes.c_int64 * MAX_LAYERS)]


_vp, _i, _f, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64, ctypes.c_size_t
_LP, _YP = ctypes.POINTER(Layout), ctypes.POINTER(Layered)

SIGNATURES = {
    "rod_last_error": (ctypes.c_char_p, []),
    "rod_version": (_i, []),
    "rod_device_info": (_i, [_vp, _vp, _vp, _vp]),
    "rod_anchor_table": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "rod_anchor_table_from_grid": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_arm_match_encode": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_arm_forced_match_workspace_bytes": (_sz, [_i, _i]),
    "rod_arm_forced_match": (_i, [_LP, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rod_odm_target": (_i, [_LP, _vp, _vp, _YP, _YP, _YP, _YP, _YP, _i, _vp, _vp, _vp, _vp, _vp]),
    "rod_target_fused_workspace_bytes": (_sz, []),
    "rod_target_fused": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _YP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_dl_target_fused": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_decode": (_i, [_LP, _vp, _YP, _YP, _i, _i, _vp, _vp]),
    "rod_decode_cascade": (_i, [_LP, _vp, _YP, _YP, _i, _i, _vp, _vp]),
    "rod_encode_one_box": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "rod_center_to_corner": (_i, [_vp, _vp, _i64, _vp]),
    "rod_corner_to_center": (_i, [_vp, _vp, _i64, _vp]),
    "rod_jaccard": (_i, [_vp, _vp, _i, _vp, _i64, _vp]),
    "rod_bboxes_jaccard": (_i, [_vp, _i, _vp, _vp, _i64, _vp]),
    "rod_bboxes_intersection": (_i, [_vp, _i, _vp, _vp, _i64, _vp]),
    "rod_bboxes_clip": (_i, [_vp, _i, _vp, _vp, _i64, _vp]),
    "rod_bboxes_resize": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "rod_bboxes_select": (_i, [_LP, _YP, _YP, _i, _i, _i, _f, _vp, _vp, _vp]),
    "rod_bboxes_sort": (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp]),
    "rod_bboxes_nms_batch": (_i, [_vp, _vp, _i64, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "rod_bboxes_matching_batch": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "rod_detect_workspace_bytes": (_sz, [_LP, _i, _i, _i]),
    "rod_detect_flags_offset": (_sz, [_LP, _i, _i, _i]),
    "rod_detect_workspace_clean_bytes": (_sz, [_LP, _i, _i, _i]),
    "rod_detect": (_i, [_LP, _vp, _YP, _YP, _YP, _YP, _i, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rod_peak_fp32_nofma": (_i, [_i, _vp, _vp, _vp]),
    "rod_l2_flush": (_i, [_vp, _sz, _vp]),
    "rod_dl_layered": (_i, [_LP, _vp, _i, _i, _i, _YP, _vp]),
    "rod_dl_arm_match_encode": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_dl_odm_target": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_dl_decode": (_i, [_LP, _vp, _vp, _vp, _i, _vp, _vp]),
    "rod_dl_detect": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_softmax": (_i, [_vp, _i64, _i, _vp, _vp]),
    "rod_detect_logits": (_i, [_LP, _vp, _YP, _YP, _YP, _YP, _i, _i, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rod_dl_detect_logits": (_i, [_LP, _vp, _vp, _vp, _vp, _vp, _i, _f, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_dl_softmax": (_i, [_vp, _vp, _vp]),
    "rod_tpfp_append": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _i, _i, _f, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "rod_sort_scores_workspace_bytes": (_sz, [_i64]),
    "rod_sort_scores_desc": (_i, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rod_precision_recall_workspace_bytes": (_sz, [_i64]),
    "rod_precision_recall": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rod_average_precision": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rod_smooth_l1_workspace_bytes": (_sz, [_LP, _i]),
    "rod_smooth_l1_loss": (_i, [_LP, _YP, _YP, _YP, _i, _vp, _vp, _f, _vp, _sz, _vp]),
    "rod_clf_loss_workspace_bytes": (_sz, [_LP, _i]),
    "rod_clf_loss": (_i, [_LP, _YP, _YP, _YP, _YP, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "rod_clf_loss_grad": (_i, [_LP, _YP, _YP, _YP, _YP, _i, _i, _f, _vp, _vp, _vp]),
    "rod_tfrecord_index": (_i, [_vp, _sz, _i, _vp, _vp]),
    "rod_tfrecord_read_gt": (_i, [_vp, _sz, _i, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rod_gt_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rod_gt_boxes_update": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _f, _i, _i, _vp, _vp, _vp, _vp]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)       # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args

E_INVALID, E_CUDA, E_UNSUPPORTED, E_DLPACK = -1, -2, -3, -4


def check(rc: int) -> None:
    """Map a C error code to the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib.rod_last_error().decode("utf-8", "replace")
    if rc in (E_INVALID, E_UNSUPPORTED, E_DLPACK):
        raise ValueError(msg)
    raise RuntimeError("rodet_b200: %s (code %d)" % (msg, rc))


_capsule_ptr = ctypes.pythonapi.PyCapsule_GetPointer
_capsule_ptr.restype = ctypes.c_void_p
_capsule_ptr.argtypes = [ctypes.py_object, ctypes.c_char_p]


class DLArgs:
    """Keeps DLPack capsules alive for the duration of one C call.  The capsules are never
    consumed (name stays "dltensor"), so dropping them releases the borrowed tensors."""

    def __init__(self):
        self._keep = []

    def one(self, t):
        if t is None:
            return None
        cap = to_dlpack(t)
        self._keep.append(cap)
        return _capsule_ptr(cap, b"dltensor")      # DLManagedTensor* == DLTensor*

    def many(self, ts):
        if ts is None:
            return None
        arr = (ctypes.c_void_p * len(ts))(*[self.one(t) for t in ts])
        self._keep.append(arr)
        return ctypes.cast(arr, ctypes.c_void_p)


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor: rodet_b200 has no CPU path" % name)
    return t


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class device_guard:
    """`torch.cuda.device(dev)` only when `dev` is not already current (the common case is free)."""
    __slots__ = ("_ctx",)

    def __init__(self, device):
        self._ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


class LayerList(list):
    """The reference's "list over layers" backed by ONE flat [B, N(, inner)] tensor.

    A real, eagerly filled `list` of the six per-layer views (so `torch.cat(lst, 1)`, `torch.stack`,
    `lst + other` and every other consumer that reads the list storage directly see the tensors, like
    the plain list the reference returns), which also remembers the flat tensor so that our own functions
    can hand it on without re-deriving six pointers (`refine_groundtruth` -> `det_groundtruth`)."""

    def __init__(self, flat, table, batched, trailing_one):
        super().__init__(table.split(flat, batched, trailing_one))
        self.flat, self.table, self.batched, self.trailing_one = flat, table, batched, trailing_one


_DT = {torch.float32: (2, 32), torch.int32: (0, 32), torch.int64: (0, 64)}


def layered_arg(ts, table, inner, dtype, args, batch_box):
    """rod_layered_t for a per-layer list.  Our own LayerList (same table) is described from its
    flat tensor directly; anything else goes through DLPack and is validated in C."""
    out = Layered()
    if (isinstance(ts, LayerList) and ts.table is table and ts.flat.dtype == dtype and ts.batched and
            len(ts) == table.n_layers and ts[0].data_ptr() == ts.flat.data_ptr() and
            ts[-1].data_ptr() == ts.flat.data_ptr() + table.offsets[-2] * inner * ts.flat.element_size()):
        flat = ts.flat                                 # (still the views it was built with: the list is mutable)
        esz = flat.element_size()
        base, stride = flat.data_ptr(), flat.stride(0)
        for l in range(table.n_layers):
            out.base[l] = base + table.offsets[l] * inner * esz
            out.batch_stride[l] = stride
        if batch_box[0] < 0:
            batch_box[0] = flat.shape[0]
        elif batch_box[0] != flat.shape[0]:
            raise ValueError("per-layer lists disagree on the batch size")
        args._keep.append(flat)
        return out
    if len(ts) != table.n_layers:
        raise ValueError("list has %d layers, anchors have %d" % (len(ts), table.n_layers))
    code, bits = _DT[dtype]
    b = ctypes.c_int(batch_box[0])
    check(lib.rod_dl_layered(table.layout, args.many(list(ts)), inner, code, bits, ctypes.byref(out), ctypes.byref(b)))
    batch_box[0] = b.value
    return out


def float_array(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def layered(ts, inner):
    """rod_layered_t for a list of per-layer tensors [B, ..., inner] (raw-pointer ABI)."""
    out = Layered()
    keep = []
    for l, t in enumerate(ts):
        if t.dim() < 2:
            raise ValueError("per-layer tensors need a batch dimension")
        per = t[0].numel()
        if not t[0].is_contiguous():
            t = t.contiguous()
        keep.append(t)
        out.base[l] = t.data_ptr()
        out.batch_stride[l] = t.stride(0) if t.shape[0] > 1 else per
    return out, keep
