#!/usr/bin/env python
"""Per-kernel timing harness (development tool; bench.py is the contract benchmark).

Times, single stream, CUDA events over graph replays on rotating input sets (working set > L2):
  arm / odm / arm+odm / fused target kernel (several tiles-per-CTA settings), B = 32
  decode_nms on the normal, stress, clustered-quadrant and clustered-bumps workloads, B = 64, with the
  share of (class, image) segments that were handed to the exact general kernels.
Usage: python profiles/bench_kernels.py [targets] [detect] [--iters N]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN  # noqa: E402
from rodet_b200 import _abi, config, synth  # noqa: E402
from rodet_b200.anchor_table import AnchorTable  # noqa: E402
from rodet_b200.utils import net_tools  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
ITERS = 200
for i, a in enumerate(sys.argv):
    if a == "--iters":
        ITERS = int(sys.argv[i + 1])
what = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()] or ["targets", "detect"]
config.img_size = BN.IMG
_anchors = net_tools.anchors_all_layer(BN.IMG, {"layer_%d" % (i + 1): f for i, f in enumerate(BN.FEATS)}, net_tools.init_anchor(len(BN.FEATS)))
config.img_size = (418, 418)
table = AnchorTable.from_anchors(_anchors, dev)
N = table.n
JB = config.refine_method.JACCARD_BIGGER


def time_graphs(graphs, iters=ITERS, warm=20):
    for i in range(warm):
        graphs[i % len(graphs)].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # us


def capture(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def to_dev_list(flat, tail):
    return [torch.from_numpy(a).to(dev) for a in BN.split_np(flat, BN.SHAPES, tail)]


res = {}
if "targets" in what:
    B, n_sets = 32, 8
    sets = []
    for s in range(n_sets):
        c, l, k, ro = BN.host_inputs_match(synth, s * B, B)
        sets.append({"center": torch.from_numpy(c).to(dev), "labels": torch.from_numpy(l).to(dev),
                     "counts": torch.from_numpy(k).to(dev), "ro": to_dev_list(ro, (4,))})
    arm = lambda s: net_tools.refine_groundtruth(table, s["center"], s["labels"], JB, gt_counts=s["counts"])
    for s in sets:
        s["arm_out"] = arm(s)
    odm = lambda s, t: net_tools.det_groundtruth(s["ro"], t[0], t[1], t[2], t[3], table)
    fused = lambda s, need=True: net_tools.target_gen(table, s["center"], s["labels"], s["ro"], gt_counts=s["counts"], need_cbboxes=need)
    res["arm_us"] = time_graphs([capture(lambda s=s: arm(s))[0] for s in sets])
    res["odm_us"] = time_graphs([capture(lambda s=s: odm(s, s["arm_out"]))[0] for s in sets])
    res["arm_odm_us"] = time_graphs([capture(lambda s=s: odm(s, arm(s)))[0] for s in sets])
    for s in sets:
        s["buf"] = net_tools.target_buffers(table, B, dev)
    fused_out = lambda s: net_tools.target_gen(table, s["center"], s["labels"], s["ro"], gt_counts=s["counts"], out=s["buf"])
    res["fused_us"] = time_graphs([capture(lambda s=s: fused_out(s))[0] for s in sets])
    torch.cuda.synchronize()
    assert all(int(s["buf"]["_sched"].abs().sum()) == 0 for s in sets), "scheduler counters not reset"
    res["fused_nocb_us"] = time_graphs([capture(lambda s=s: fused(s, False))[0] for s in sets])
    # parity of the fused kernel against the two-call path on every set
    same = True
    for s in sets:
        a1 = arm(s)
        d1 = odm(s, a1)
        a2, d2 = fused(s)
        same &= all(torch.equal(x.flat, y.flat) for x, y in zip(a1, a2)) and all(
            torch.equal(x.flat.view(torch.int32), y.flat.view(torch.int32)) for x, y in zip(d1, d2))
    res["fused_equals_two_calls"] = bool(same)
    bytes_m = B * (124 * N + 20 * 50.5)
    res["hbm_frac_fused"] = bytes_m / (res["fused_us"] * 1e-6) / 1e9 / 6545.6
    res["hbm_frac_arm_odm"] = bytes_m / (res["arm_odm_us"] * 1e-6) / 1e9 / 6545.6

if "detect" in what:
    B, n_sets = 64, 3
    kw = dict(select_threshold=BN.SELECT_THR, nms_threshold=BN.NMS_THR, top_k=BN.TOP_K, keep_top_k=BN.KEEP, return_counts=True)
    tag = ""
    if True:
      for name in os.environ.get("ROD_BK_KINDS", "normal,stress,quadrant,bumps").split(","):
        graphs, wss, keep = [], [], []
        for s in range(n_sets):
            first = 500_000 + s * B
            p, ro, do = BN.host_inputs_detect(synth, first, B, name)
            d = {"p": to_dev_list(p, (11,)), "ro": to_dev_list(ro, (4,)), "do": to_dev_list(do, (4,))}
            ws = net_tools.detect_workspace(table, B, BN.TOP_K, dev)
            g, out = capture(lambda d=d, ws=ws: net_tools.decode_detected_bboxes(table, d["ro"], d["do"], d["p"], workspace=ws, **kw))
            graphs.append(g)
            wss.append(ws)
            d["out"] = out
            keep.append(d)
        us = time_graphs(graphs, max(20, ITERS // 2), 6)
        print(name, tag, us, file=sys.stderr, flush=True)
        rate = float(np.mean([net_tools.detect_fallback_flags(w)[1:].float().mean().item() for w in wss]))
        res["detect_%s%s_us" % (name, tag)] = us
        res["detect_%s%s_fallback_rate" % (name, tag)] = rate
        res["detect_%s%s_hbm_frac" % (name, tag)] = B * (76 * N + 10 * BN.KEEP * 20) / (us * 1e-6) / 1e9 / 6545.6
print(json.dumps(res, indent=1))
