"""Device-resident anchor table shared by every kernel of the path.

For N anchors (flattened layer-major, then (fy, fx, a) row-major) it holds
  corner[N,4] = (ymin, xmin, ymax, xmax)   utils/net_tools.py:156-165
  center[N,4] = (acy, acx, ah, aw)         re-derived from the corners, :168-171
both float32, 16 B per anchor per table (590 KB at 512x512: L2-resident)."""
from __future__ import annotations

import ctypes
import zlib

import numpy as np
import torch

from . import _abi


class AnchorTable:
    def __init__(self, shapes, corner, center, yxhw=None):
        self.shapes = [tuple(int(v) for v in s) for s in shapes]      # (fh, fw, A) per layer
        self.counts = [fh * fw * a for fh, fw, a in self.shapes]
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)
        self.n = self.offsets[-1]
        self.n_layers = len(self.shapes)
        if self.n_layers > _abi.MAX_LAYERS:
            raise ValueError("at most %d layers are supported" % _abi.MAX_LAYERS)
        self.corner, self.center, self.yxhw = corner, center, yxhw
        self.device = corner.device
        self.layout = _abi.Layout()
        self.layout.n_layers = self.n_layers
        self.layout.n_total = self.n
        for i, o in enumerate(self.offsets):
            self.layout.offset[i] = o
        self._sub = {}

    # ------------------------------------------------------------------ constructors
    @classmethod
    def generate(cls, img_size, feat_sizes, anchor_sizes_px, device="cuda"):
        """Run a1-a4 on the device: `anchor_sizes_px` is init_anchor()'s OrderedDict (or a list
        of float64 [A,2] arrays), `feat_sizes` a dict layer_i -> (fh, fw) or a list."""
        device = torch.device(device)
        sizes = list(anchor_sizes_px.values()) if isinstance(anchor_sizes_px, dict) else list(anchor_sizes_px)
        keys = list(anchor_sizes_px.keys()) if isinstance(anchor_sizes_px, dict) else None
        feats = [feat_sizes[k] for k in keys] if isinstance(feat_sizes, dict) else list(feat_sizes)
        nl = len(sizes)
        shapes = [(int(feats[l][0]), int(feats[l][1]), int(np.asarray(sizes[l]).shape[0])) for l in range(nl)]
        n = sum(a * b * c for a, b, c in shapes)
        fh = (ctypes.c_int32 * nl)(*[s[0] for s in shapes])
        fw = (ctypes.c_int32 * nl)(*[s[1] for s in shapes])
        na = (ctypes.c_int32 * nl)(*[s[2] for s in shapes])
        px = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in sizes]), dtype=np.float64)
        corner = torch.empty((n, 4), dtype=torch.float32, device=device)
        center = torch.empty_like(corner)
        yxhw = torch.empty_like(corner)
        with torch.cuda.device(device):
            _abi.check(_abi.lib.rod_anchor_table(nl, fh, fw, na, px.ctypes.data, int(img_size[0]), int(img_size[1]),
                                                 corner.data_ptr(), center.data_ptr(), yxhw.data_ptr(),
                                                 _abi.stream_ptr(device)))
        return cls(shapes, corner, center, yxhw)

    @classmethod
    def from_anchors(cls, anchors_all_layer, device="cuda"):
        """From an `anchors_all_layer()` result: list over layers of [y[fh,fw,1], x[fh,fw,1], h[A], w[A]]."""
        device = torch.device(device)
        nl = len(anchors_all_layer)
        shapes, ys, xs, hs, ws = [], [], [], [], []
        for y, x, h, w in anchors_all_layer:
            y, x, h, w = (np.asarray(v, dtype=np.float32) for v in (y, x, h, w))
            shapes.append((y.shape[0], y.shape[1], h.shape[0]))
            ys.append(y.reshape(-1)); xs.append(x.reshape(-1)); hs.append(h.reshape(-1)); ws.append(w.reshape(-1))
        n = sum(a * b * c for a, b, c in shapes)
        up = lambda parts: torch.from_numpy(np.ascontiguousarray(np.concatenate(parts))).to(device)
        y_d, x_d, h_d, w_d = up(ys), up(xs), up(hs), up(ws)
        fh = (ctypes.c_int32 * nl)(*[s[0] for s in shapes])
        fw = (ctypes.c_int32 * nl)(*[s[1] for s in shapes])
        na = (ctypes.c_int32 * nl)(*[s[2] for s in shapes])
        corner = torch.empty((n, 4), dtype=torch.float32, device=device)
        center = torch.empty_like(corner)
        with torch.cuda.device(device):
            _abi.check(_abi.lib.rod_anchor_table_from_grid(nl, fh, fw, na, y_d.data_ptr(), x_d.data_ptr(),
                                                           h_d.data_ptr(), w_d.data_ptr(), corner.data_ptr(),
                                                           center.data_ptr(), _abi.stream_ptr(device)))
        return cls(shapes, corner, center)

    # ------------------------------------------------------------------ helpers
    def layer(self, l):
        """Single-layer sub-table (views, no copy)."""
        if l not in self._sub:
            s = slice(self.offsets[l], self.offsets[l + 1])
            self._sub[l] = AnchorTable([self.shapes[l]], self.corner[s], self.center[s])
        return self._sub[l]

    def split(self, flat, batched=True, trailing_one=False):
        """flat [B,N] or [B,N,4] -> list over layers of [B,fh,fw,A(,4)] views (batch dim dropped
        when `batched` is False; `trailing_one` appends a size-1 axis, as the reference does
        for labels and masks)."""
        out = []
        B, st, off0 = flat.shape[0], flat.stride(), flat.storage_offset()
        s1 = st[1]
        for l, (fh, fw, a) in enumerate(self.shapes):
            shape, stride = [B, fh, fw, a], [st[0], fw * a * s1, a * s1, s1]
            if flat.dim() == 3:
                shape.append(flat.shape[2]); stride.append(st[2])
            elif trailing_one:
                shape.append(1); stride.append(1)
            if not batched:
                shape, stride = shape[1:], stride[1:]
            out.append(flat.as_strided(shape, stride, off0 + self.offsets[l] * s1))
        return out


_CACHE = {}
_CACHE_MAX = 64


def _content_key(layers, device, tag):
    """Cache key from the CONTENT of the anchors (shapes + CRC32 and Adler-32 of the y / x / h / w bytes,
    ~10 us for the 512x512 layout), so that anchors mutated in place or rebuilt every step never map to a
    stale table and equal anchors rebuilt by the caller hit the cache."""
    crc, adl, shapes = 0, 1, []
    for layer in layers:
        for v in layer:
            a = np.ascontiguousarray(v, dtype=np.float32)
            shapes.append(a.shape)
            crc = zlib.crc32(a, crc)
            adl = zlib.adler32(a, adl)
    return (tag, str(device), tuple(shapes), crc, adl)


def _cached(layers, device, tag):
    key = _content_key(layers, device, tag)
    t = _CACHE.get(key)
    if t is None:
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))          # drop the oldest entry only
        t = _CACHE[key] = AnchorTable.from_anchors(layers, device)
    return t


def table_for(anchors, device):
    """AnchorTable for an anchors_all_layer() list, cached on the anchors' content.  Hot loops should
    build the AnchorTable once and pass it instead of the list (skips the hashing)."""
    if isinstance(anchors, AnchorTable):
        return anchors
    return _cached(anchors, device, "all")


def layer_table_for(anchors_one_layer, device):
    """AnchorTable for ONE layer's [y, x, h, w] (what decode/encode_locations_one_layer take)."""
    if isinstance(anchors_one_layer, AnchorTable):
        return anchors_one_layer
    return _cached([anchors_one_layer], device, "one")
