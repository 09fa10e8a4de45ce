"""utils/tf_extended/tensors.py:34-86 — shape / padding helpers (pure tensor plumbing)."""
import torch

__all__ = ["get_shape", "pad_axis"]


def get_shape(x, rank=None):
    """Dimensions of `x` as a list of ints (torch shapes are always static)."""
    shape = list(x.shape)
    if rank is not None and len(shape) != rank:
        raise ValueError("expected rank %d, got %d" % (rank, len(shape)))
    return shape


def pad_axis(x, offset, size, axis=0, name=None):
    """Zero-pad `x` on `axis` with `offset` leading zeros up to `size` entries; never
    truncates (utils/tf_extended/tensors.py:59-86)."""
    n = x.shape[axis]
    after = max(size - offset - n, 0)
    if offset == 0 and after == 0:
        return x
    shape = list(x.shape)
    shape[axis] = offset + n + after
    out = x.new_zeros(shape)
    out.narrow(axis, offset, n).copy_(x)
    return out
