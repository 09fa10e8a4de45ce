#!/usr/bin/env python
"""Timing of the f-3 path (development tool): target_gen + refine_loss + det_clf_loss at B = 32, 512 x 512.
  graph   : forward captured into a CUDA graph (8 steps per graph), ms per step
  eager   : the same forward launched through Python, ms per step
  losses  : the three losses alone (targets precomputed), graph, ms per step
  fwd+bwd : eager forward + backward"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as BN                                       # noqa: E402
from rodet_b200 import config, synth                     # noqa: E402
from rodet_b200.anchor_table import AnchorTable          # noqa: E402
from rodet_b200.utils import net_tools                   # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
config.img_size = BN.IMG
anchors = net_tools.anchors_all_layer(BN.IMG, {"layer_%d" % (i + 1): f for i, f in enumerate(BN.FEATS)}, net_tools.init_anchor(6))
config.img_size = (418, 418)
table = AnchorTable.from_anchors(anchors, dev)
N, B = table.n, 32
lay = lambda flat, tail: [torch.from_numpy(a).to(dev) for a in BN.split_np(flat, BN.SHAPES, tail)]
c, l, k, ro = BN.host_inputs_match(synth, 800_000, B)
cen, lab, cnt = torch.from_numpy(c).to(dev), torch.from_numpy(l).to(dev), torch.from_numpy(k).to(dev)
ro_l = lay(ro, (4,))
do_l = lay(np.stack([synth.head_offsets(800_000 + b, N, 1) for b in range(B)]), (4,))
clf_l = lay(np.stack([synth.class_logits(800_000 + b, N) for b in range(B)]), (11,))
buf = net_tools.target_buffers(table, B, dev)


def losses(ro_, do_, clf_, arm, det):
    rl = net_tools.refine_loss(ro_, arm[0], arm[3])
    dl, cl = net_tools.det_clf_loss(ro_, clf_, do_, det[0], det[1], det[2], det[3])
    return rl + dl + cl


def forward(ro_, do_, clf_):
    arm, det = net_tools.target_gen(table, cen, lab, [t.detach() for t in ro_], gt_counts=cnt, out=buf)
    return losses(ro_, do_, clf_, arm, det)


def timed(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graph_of(fn, K=8):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(K):
            fn()
    return g, K


res, keep = {}, {}
with torch.no_grad():
    res["loss"] = float(forward(ro_l, do_l, clf_l))
    g, K = graph_of(lambda: keep.__setitem__("l", forward(ro_l, do_l, clf_l)))
    res["forward_graph_ms"] = timed(g.replay, 50) / K
    res["forward_eager_ms"] = timed(lambda: forward(ro_l, do_l, clf_l), 50)
    arm, det = net_tools.target_gen(table, cen, lab, ro_l, gt_counts=cnt, out=buf)
    g2, K = graph_of(lambda: keep.__setitem__("l2", losses(ro_l, do_l, clf_l, arm, det)))
    res["losses_graph_ms"] = timed(g2.replay, 50) / K
    res["losses_eager_ms"] = timed(lambda: losses(ro_l, do_l, clf_l, arm, det), 50)
leaves = [[x.clone().requires_grad_(True) for x in ts] for ts in (ro_l, do_l, clf_l)]


def step_bwd():
    for ts in leaves:
        for x in ts:
            x.grad = None
    forward(*leaves).backward()


res["forward_backward_eager_ms"] = timed(step_bwd, 30)
print(json.dumps(res))
