"""Shared helpers for the GPU parity tests (torch <-> oracle plumbing)."""
import numpy as np
import torch

from oracle import restated as R


def split_layers_np(flat, shapes, tail):
    """[B,N,*tail] -> list of [B,fh,fw,A,*tail] numpy arrays."""
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        out.append(np.ascontiguousarray(flat[:, off:off + n]).reshape((flat.shape[0], fh, fw, a) + tuple(tail)))
        off += n
    return out


def to_cuda_list(flat, shapes, tail, dev):
    return [torch.from_numpy(a).to(dev) for a in split_layers_np(flat, shapes, tail)]


def flat_from_list(ts, tail_dims):
    """list of [B,fh,fw,A,*tail] torch tensors -> numpy [B,N,*tail]."""
    outs = []
    for t in ts:
        a = t.detach().cpu().numpy()
        if tail_dims:
            outs.append(a.reshape(a.shape[0], -1, *a.shape[a.ndim - tail_dims:]))
        else:
            outs.append(a.reshape(a.shape[0], -1))
    return np.concatenate(outs, axis=1)


def bit_equal(a, b):
    """Exact equality for float arrays where -0.0 == +0.0 and NaN == NaN position-wise."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def oracle_table(anchors):
    return R.AnchorTable(anchors)
