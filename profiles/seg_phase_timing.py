#!/usr/bin/env python
"""Per-phase clock64() breakdown of segment_kernel (debug hook rod_debug_set_timing)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rodet_b200 import _abi, config, synth  # noqa: E402
from rodet_b200.anchor_table import AnchorTable  # noqa: E402
from rodet_b200.utils import net_tools  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "normal"
dev = torch.device("cuda:0")
config.img_size = bench.IMG
anchors = net_tools.anchors_all_layer(bench.IMG, {"layer_%d" % (i + 1): f for i, f in enumerate(bench.FEATS)}, net_tools.init_anchor(6))
config.img_size = (418, 418)
table = AnchorTable.from_anchors(anchors, dev)
B = 64
p, ro, do = bench.host_inputs_detect(synth, 500_000, B, kind)
tl = lambda a, t: [torch.from_numpy(x).to(dev) for x in bench.split_np(a, bench.SHAPES, t)]
P, RO, DO = tl(p, (11,)), tl(ro, (4,)), tl(do, (4,))
dbg = torch.zeros((2 * 11 * B, 8), dtype=torch.int64, device=dev)
_abi.lib.rod_debug_set_timing.argtypes = [ctypes.c_void_p]
run = lambda: net_tools.decode_detected_bboxes(table, RO, DO, P, select_threshold=0.3, nms_threshold=0.45, top_k=400, keep_top_k=200)
for _ in range(3):
    run()
_abi.lib.rod_debug_set_timing(dbg.data_ptr())
run()
torch.cuda.synchronize()
_abi.lib.rod_debug_set_timing(None)
allrows = dbg.cpu().numpy()
d = allrows[B:11 * B]                          # class 0 rows are skipped

t = d[:, :6].astype(np.float64)
names = ["load+filter", "sort", "gather/decode", "nms", "emit"]
dur = np.diff(t, axis=1)
smid = (d[:, 7] >> 32).astype(np.int64)
d[:, 7] &= 0xffffffff
print("segments", len(d), "mean kept", d[:, 6].mean(), "mean list", d[:, 7].mean())
for n, v in zip(names, dur.mean(0)):
    print("%-14s %8.0f cycles  %6.2f us @1.9GHz" % (n, v, v / 1900.0))
print("total          %8.0f cycles  %6.2f us ; start spread %.1f us, end spread %.1f us" % (
    (t[:, 5] - t[:, 0]).mean(), (t[:, 5] - t[:, 0]).mean() / 1900.0,
    (t[:, 0].max() - t[:, 0].min()) / 1900.0, (t[:, 5].max() - t[:, 5].min()) / 1900.0))

tot = (t[:, 5] - t[:, 0]) / 1900.0
print("total per CTA, us: min %.1f  p50 %.1f  p90 %.1f  p99 %.1f  max %.1f" % (tot.min(), *np.percentile(tot, [50, 90, 99]), tot.max()))
for n, v in zip(names, dur.T):
    v = v / 1900.0
    print("%-14s p50 %6.2f  p90 %6.2f  p99 %6.2f  max %6.2f" % (n, *np.percentile(v, [50, 90, 99]), v.max()))
slow = np.argsort(tot)[-10:]
print("ten slowest CTAs: total / load+filter / nms (us):", [(round(float(tot[i]), 1), round(float(dur[i, 0] / 1900.0), 1), round(float(dur[i, 3] / 1900.0), 1)) for i in slow])
per_sm = np.bincount(smid, minlength=148)
print("segments per SM: histogram", dict(zip(*np.unique(per_sm, return_counts=True))))
for c in np.unique(per_sm):
    sel = np.isin(smid, np.nonzero(per_sm == c)[0])
    if sel.any():
        print("  SMs holding %d segments: mean CTA time %.1f us, max %.1f" % (c, tot[sel].mean(), tot[sel].max()))
