"""Ground truth from the reference's on-disk format.

The reference writes one `tf.train.Example` per image into TFRecord files (dataset/pascalvoc_to_tfrecords.py:128-170)
and reads them back with slim's `DatasetDataProvider` over the schema of dataset/pascalvoc_common.py:75-98
(`get_split`), feeding `object/bbox` ([ymin, xmin, ymax, xmax]), `object/label` and `object/difficult` to
prepare_data_train / prepare_data_test (utils/data_pileline_tools.py:33-71).  Here the box-level fields of ALL records
are parsed once by the native reader (csrc/tfrecord.cu), uploaded to the GPU once, and every batch is assembled on
the device (`rod_gt_gather`) into the padded form the rest of the path takes:

    gt = read_ground_truth(glob('bdd100k_train_*.tfrecord'))           # host, once
    dgt = gt.to('cuda')                                                # HBM, once (BDD100K: ~40 MB)
    bboxes, labels, difficults, counts = dgt.batch(indices)            # device, per step
    center = cornerBboxes_2_centerBboxes(bboxes); refine_groundtruth(anchors, center, labels, ..., gt_counts=counts)

Image bytes ('image/encoded') are skipped: JPEG decoding and augmentation are outside the box-level path."""
from __future__ import annotations

import ctypes
import mmap
import os

import numpy as np
import torch

from .. import _abi

# the keys this reader consumes (dataset/pascalvoc_common.py:82-97)
BBOX_KEYS = ('image/object/bbox/ymin', 'image/object/bbox/xmin', 'image/object/bbox/ymax', 'image/object/bbox/xmax')
LABEL_KEY, DIFFICULT_KEY, TRUNCATED_KEY, SHAPE_KEY = ('image/object/bbox/label', 'image/object/bbox/difficult',
                                                      'image/object/bbox/truncated', 'image/shape')


class GroundTruth:
    """Ragged ground truth of a dataset: per object ymin/xmin/ymax/xmax (float32), label/difficult/truncated (int64);
    `offsets[r] : offsets[r+1]` are the objects of record r; `shape[r]` = (height, width, channels)."""

    def __init__(self, ymin, xmin, ymax, xmax, label, difficult, truncated, offsets, shape):
        self.ymin, self.xmin, self.ymax, self.xmax = ymin, xmin, ymax, xmax
        self.label, self.difficult, self.truncated, self.offsets, self.shape = label, difficult, truncated, offsets, shape

    def __len__(self):
        return int(self.offsets.shape[0]) - 1

    @property
    def max_objects(self):
        d = np.diff(self.offsets) if isinstance(self.offsets, np.ndarray) else (self.offsets[1:] - self.offsets[:-1])
        return int(d.max()) if len(self) else 0

    def to(self, device):
        """Upload once; `batch()` then runs entirely on the device."""
        dev = torch.device(device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if isinstance(a, np.ndarray) else a.to(dev)
        g = GroundTruth(*[t(a) for a in (self.ymin, self.xmin, self.ymax, self.xmax, self.label, self.difficult,
                                         self.truncated, self.offsets, self.shape)])
        g._max = self.max_objects
        return g

    def batch(self, indices=None, max_gt=None, batch=None):
        """Padded batch on the device: (bboxes [B,G,4] float32 corner form, labels [B,G] int64, difficults [B,G] int64,
        counts [B] int32).  `indices`: sequence / tensor of record numbers (None: records 0..batch-1; host-side indices are
        range-checked here, a CUDA tensor is used as it is and marks invalid entries with counts = -1);
        G = max_gt (default: the largest object count of the dataset); rows are zero padded, counts clipped to G."""
        if not isinstance(self.offsets, torch.Tensor) or not self.offsets.is_cuda:
            raise ValueError("call .to('cuda') first: batches are assembled on the device (there is no CPU path)")
        dev = self.offsets.device
        if indices is not None:
            if isinstance(indices, torch.Tensor) and indices.is_cuda:
                # device-resident indices are not read back (no round trip per step): an index outside the dataset
                # gives an all-zero row with counts = -1
                idx = indices.to(device=dev, dtype=torch.int64).contiguous()
            else:
                host = np.asarray(indices.cpu() if isinstance(indices, torch.Tensor) else indices, dtype=np.int64).reshape(-1)
                if host.size and (int(host.min()) < 0 or int(host.max()) >= len(self)):
                    raise IndexError("record index out of range")
                idx = torch.from_numpy(host).to(dev)
            B = int(idx.numel())
        else:
            idx, B = None, len(self) if batch is None else int(batch)
            if B > len(self):
                raise IndexError("batch exceeds the number of records")
        G = int(max_gt) if max_gt is not None else max(1, getattr(self, "_max", 1))
        bboxes = torch.empty((B, G, 4), dtype=torch.float32, device=dev)
        labels = torch.empty((B, G), dtype=torch.int64, device=dev)
        diff = torch.empty((B, G), dtype=torch.int64, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        P = lambda t: t.data_ptr() if t is not None and t.numel() else None
        with _abi.device_guard(dev):
            _abi.check(_abi.lib.rod_gt_gather(P(self.ymin), P(self.xmin), P(self.ymax), P(self.xmax), P(self.label),
                                              P(self.difficult), self.offsets.data_ptr(), P(idx), len(self), B, G, bboxes.data_ptr(),
                                              labels.data_ptr(), diff.data_ptr(), counts.data_ptr(), _abi.stream_ptr(dev)))
        return bboxes, labels, diff, counts


def read_ground_truth(files, verify_crc=True):
    """Parses the ground-truth features of every record of one or more TFRecord files (paths, or bytes objects) with the
    native reader.  Raises ValueError on corrupted / malformed records.  Returns a host `GroundTruth` (NumPy arrays)."""
    if isinstance(files, (str, bytes, bytearray, memoryview)):
        files = [files]
    parts = []
    for f in files:                             # one file at a time (the records carry the JPEG bytes too): mapped, not copied
        if isinstance(f, str):
            if os.path.getsize(f) == 0:
                continue
            with open(f, "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
                parts.append(_parse(np.frombuffer(mm, dtype=np.uint8), verify_crc))
        else:
            parts.append(_parse(np.frombuffer(bytes(f), dtype=np.uint8), verify_crc))
    if not parts:
        parts.append(_parse(np.zeros(0, dtype=np.uint8), verify_crc))
    cat = lambda i: np.concatenate([p[i] for p in parts])
    offsets, base = [np.zeros(1, dtype=np.int64)], 0
    for p in parts:                             # TFRecord files concatenate: shift every file's offsets
        offsets.append(p[7][1:] + base)
        base += int(p[7][-1])
    return GroundTruth(cat(0), cat(1), cat(2), cat(3), cat(4), cat(5), cat(6), np.concatenate(offsets), cat(8))


def _parse(data, verify_crc):
    """One file's bytes (uint8 array, may be a read-only mapping) -> the nine arrays of GroundTruth."""
    ptr = data.ctypes.data if data.size else None
    n_rec, n_obj = ctypes.c_int64(0), ctypes.c_int64(0)
    _abi.check(_abi.lib.rod_tfrecord_index(ptr, data.size, 1 if verify_crc else 0, ctypes.byref(n_rec), ctypes.byref(n_obj)))
    R, O = n_rec.value, n_obj.value
    f32 = lambda: np.zeros(O, dtype=np.float32)
    i64 = lambda: np.zeros(O, dtype=np.int64)
    ymin, xmin, ymax, xmax, label, difficult, truncated = f32(), f32(), f32(), f32(), i64(), i64(), i64()
    offsets, shape = np.zeros(R + 1, dtype=np.int64), np.zeros((R, 3), dtype=np.int64)
    p = lambda a: a.ctypes.data if a.size else None
    _abi.check(_abi.lib.rod_tfrecord_read_gt(ptr, data.size, 0, R, O, p(ymin), p(xmin), p(ymax), p(xmax), p(label), p(difficult),
                                             p(truncated), offsets.ctypes.data, p(shape)))
    return ymin, xmin, ymax, xmax, label, difficult, truncated, offsets, shape
