"""Runs ARM + ODM + refine_loss + det_clf_loss (+ backward) once or twice at B = 32, 512x512, for ncu launch lists:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/loss_launches.csv python profiles/run_losses_once.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rodet_b200 import config, synth                    # noqa: E402
from rodet_b200.utils import net_tools                  # noqa: E402

dev = torch.device("cuda:0")
B, img = 32, (512, 512)
feats = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
config.img_size = img
anchors = net_tools.anchors_all_layer(img, {"layer_%d" % (i + 1): f for i, f in enumerate(feats)}, net_tools.init_anchor(6))
shapes = [(fh, fw, a) for (fh, fw), a in zip(feats, net_tools.n_anchor_each_layer("mobilenet_v2"))]
N = sum(h * w * a for h, w, a in shapes)


def layers(flat, tail):
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        out.append(torch.from_numpy(np.ascontiguousarray(flat[:, off:off + n]).reshape((B, fh, fw, a) + tail)).to(dev).requires_grad_(True))
        off += n
    return out


corner, labels, counts = synth.gt_batch(0, B)
center = np.stack([(corner[..., 0] + corner[..., 2]) / 2, (corner[..., 1] + corner[..., 3]) / 2, corner[..., 2] - corner[..., 0],
                   corner[..., 3] - corner[..., 1]], -1).astype(np.float32)
ro = layers(np.stack([synth.head_offsets(b, N) for b in range(B)]), (4,))
do = layers(np.stack([synth.head_offsets(b, N, 1) for b in range(B)]), (4,))
clf = layers(np.stack([synth.class_logits(b, N) for b in range(B)]), (11,))
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for it in range(2):
    gt, cb, lab, pos = net_tools.refine_groundtruth(anchors, d(center), d(labels), config.refine_method.JACCARD_BIGGER, gt_counts=d(counts))
    det_gt, mask, dlab, iou = net_tools.det_groundtruth([t.detach() for t in ro], gt, cb, lab, pos, anchors)
    rl = net_tools.refine_loss(ro, gt, pos)
    dl, cl = net_tools.det_clf_loss(ro, clf, do, det_gt, mask, dlab, iou)
    (rl + dl + cl).backward()
torch.cuda.synchronize()
print("losses:", float(rl), float(dl), float(cl))
