"""Box format helpers with the reference's names (utils/common_tools.py:16-57), on CUDA
torch tensors of any leading shape with a last dimension of 4."""
from __future__ import annotations

import logging

import torch

from .. import _abi

logger = logging.getLogger(__name__)


def _boxop(fn, boxes, name):
    t = _abi.require_cuda(boxes, name)
    if t.dtype != torch.float32 or t.shape[-1] != 4:
        raise ValueError("%s: expected a float32 tensor with last dimension 4" % name)
    src = t.contiguous()
    out = torch.empty_like(src)
    with torch.cuda.device(t.device):
        _abi.check(fn(src.data_ptr(), out.data_ptr(), src.numel() // 4, _abi.stream_ptr(t.device)))
    return out


def centerBboxes_2_cornerBboxes(center_bboxes):
    """[yc, xc, h, w] -> [ymin, xmin, ymax, xmax] (utils/common_tools.py:16-35)."""
    return _boxop(_abi.lib.rod_center_to_corner, center_bboxes, "center_bboxes")


def cornerBboxes_2_centerBboxes(corner_bboxes):
    """[ymin, xmin, ymax, xmax] -> [yc, xc, h, w] (utils/common_tools.py:38-57)."""
    return _boxop(_abi.lib.rod_corner_to_center, corner_bboxes, "corner_bboxes")
