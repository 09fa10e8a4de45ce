"""Runs the decode_nms workload (B = 64, 512x512) a few times, from probabilities and from logits, so that
ncu can capture scan_kernel<*,11,false|true> / segment_kernel in isolation:
  ncu --set full --import-source on --clock-control none -k regex:"^(scan_kernel|segment_kernel)$" -c 8 \
      -o gpurun_out/detect python profiles/run_detect_once.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rodet_b200 import config, synth                    # noqa: E402
from rodet_b200.utils import net_tools                  # noqa: E402

dev = torch.device("cuda:0")
B, img = 64, (512, 512)
feats = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
config.img_size = img
anchors = net_tools.anchors_all_layer(img, {"layer_%d" % (i + 1): f for i, f in enumerate(feats)}, net_tools.init_anchor(6))
shapes = [(fh, fw, a) for (fh, fw), a in zip(feats, net_tools.n_anchor_each_layer("mobilenet_v2"))]
N = sum(h * w * a for h, w, a in shapes)


def layers(flat, tail):
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        out.append(torch.from_numpy(np.ascontiguousarray(flat[:, off:off + n]).reshape((B, fh, fw, a) + tail)).to(dev))
        off += n
    return out


z = np.stack([synth.class_logits(b, N) for b in range(B)])
p = np.stack([synth.class_probs(b, N) for b in range(B)])
ro = layers(np.stack([synth.head_offsets(b, N, 0, 0.1, 0.2) for b in range(B)]), (4,))
do = layers(np.stack([synth.head_offsets(b, N, 1, 0.1, 0.2) for b in range(B)]), (4,))
zl, pl = layers(z, (11,)), layers(p, (11,))
kw = dict(select_threshold=0.3, nms_threshold=0.45, top_k=400, keep_top_k=200)
for it in range(2):
    a = net_tools.decode_detected_bboxes(anchors, ro, do, pl, **kw)
    b = net_tools.decode_detected_bboxes(anchors, ro, do, zl, from_logits=True, **kw)
torch.cuda.synchronize()
print("detections:", sum(int((a[0][c] != 0).sum()) for c in a[0]), sum(int((b[0][c] != 0).sum()) for c in b[0]))
