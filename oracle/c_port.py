"""ctypes wrapper of the C oracle (oracle/c/rodet_oracle.c).  TEST INFRASTRUCTURE ONLY.

Used by tests as a fast checker for full-size runs and by bench.py as the timed multi-threaded
CPU baseline / reference arm.  Built by oracle/c/Makefile (gcc -O2 -fopenmp -ffp-contract=off)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "librodet_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, stdout=subprocess.DEVNULL)
        _lib = ctypes.CDLL(LIB)
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def threads():
    return int(load().orc_max_threads())


def set_threads(n):
    load().orc_set_threads(int(n))


def per_anchor_thresholds(table, thresholds):
    return np.ascontiguousarray(np.asarray(thresholds, dtype=np.float32)[table.layer_of])


def arm_match_encode(table, boxes, labels, counts, thresholds):
    """boxes [B,gmax,4] centre form, labels [B,gmax] int64, counts [B] -> flat gt, cb, lab, pos, idx."""
    lib = load()
    B, gmax = boxes.shape[:2]
    n = table.n
    thr = per_anchor_thresholds(table, thresholds)
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    gt = np.empty((B, n, 4), np.float32); cb = np.empty((B, n, 4), np.float32)
    lab = np.empty((B, n), np.int32); pos = np.empty((B, n), np.int32); idx = np.empty((B, n), np.int32)
    lib.orc_arm_match_encode(_p(table.corner), _p(table.center), _p(thr), n, _p(boxes), _p(labels), _p(counts),
                             B, gmax, _p(gt), _p(cb), _p(lab), _p(pos), _p(idx))
    return gt, cb, lab, pos, idx


def odm_target(table, refine_out, offset_gt, cbboxes, labels, pos, thresholds):
    lib = load()
    B, n = refine_out.shape[:2]
    thr = per_anchor_thresholds(table, thresholds)
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    ro, og, cbb = c(refine_out, np.float32), c(offset_gt, np.float32), c(cbboxes, np.float32)
    lb, pm = c(labels, np.int32), c(pos, np.int32)
    det_gt = np.empty((B, n, 4), np.float32); mask = np.empty((B, n), np.int32)
    dl = np.empty((B, n), np.int32); iou = np.empty((B, n), np.float32)
    lib.orc_odm_target(_p(table.center), _p(thr), n, B, _p(ro), _p(og), _p(cbb), _p(lb), _p(pm), _p(det_gt),
                       _p(mask), _p(dl), _p(iou))
    return det_gt, mask, dl, iou


def decode_corner(table, refine_out, det_out):
    lib = load()
    B, n = refine_out.shape[:2]
    ro = np.ascontiguousarray(refine_out, dtype=np.float32); do = np.ascontiguousarray(det_out, dtype=np.float32)
    out = np.empty((B, n, 4), np.float32)
    lib.orc_decode_corner(_p(table.center), n, B, _p(ro), _p(do), _p(out))
    return out


def detected_bboxes(probs, boxes, select_threshold, nms_threshold, top_k, keep_top_k):
    """probs [B,N,C], boxes [B,N,4] -> dicts c -> [B,keep], c -> [B,keep,4] for c = 1..C-1."""
    lib = load()
    B, n, C = probs.shape
    probs = np.ascontiguousarray(probs, dtype=np.float32); boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    thr = 0.0 if select_threshold is None else select_threshold
    os_ = np.zeros((C, B, keep_top_k), np.float32); ob = np.zeros((C, B, keep_top_k, 4), np.float32)
    lib.orc_detect(_p(probs), _p(boxes), B, n, C, ctypes.c_float(thr), ctypes.c_float(nms_threshold), int(top_k),
                   int(keep_top_k), _p(os_), _p(ob))
    return {c: os_[c] for c in range(1, C)}, {c: ob[c] for c in range(1, C)}
