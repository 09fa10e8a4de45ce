"""Runs every hot-path kernel a few times for ncu captures (round 2):
  fused target generation (B = 32), decode_nms (B = 64) on the normal / stress / bumps-clustered workloads.
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python profiles/run_once.py
  ncu --set full --import-source on --clock-control none -k regex:"target_fused|sample_kernel|scan_kernel|segment_kernel" -s 8 -c 8 -o gpurun_out/full python profiles/run_once.py normal"""
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
kinds = [a for a in sys.argv[1:]] or ["normal", "stress", "bumps"]
config.img_size = BN.IMG
anchors = net_tools.anchors_all_layer(BN.IMG, {"layer_%d" % (i + 1): f for i, f in enumerate(BN.FEATS)}, net_tools.init_anchor(6))
config.img_size = (418, 418)
table = AnchorTable.from_anchors(anchors, dev)
N = table.n
lay = lambda flat, tail: [torch.from_numpy(a).to(dev) for a in BN.split_np(flat, BN.SHAPES, tail)]

B = 32
c, l, k, ro = BN.host_inputs_match(synth, 0, B)
cen, lab, cnt, ro_l = torch.from_numpy(c).to(dev), torch.from_numpy(l).to(dev), torch.from_numpy(k).to(dev), lay(ro, (4,))
buf = net_tools.target_buffers(table, B, dev)
for it in range(3):
    net_tools.target_gen(table, cen, lab, ro_l, gt_counts=cnt, out=buf)
    t = net_tools.refine_groundtruth(table, cen, lab, config.refine_method.JACCARD_BIGGER, gt_counts=cnt)
    net_tools.det_groundtruth(ro_l, t[0], t[1], t[2], t[3], table)
torch.cuda.synchronize()

B = 64
kw = dict(select_threshold=BN.SELECT_THR, nms_threshold=BN.NMS_THR, top_k=BN.TOP_K, keep_top_k=BN.KEEP)
for kind in kinds:
    p, ro, do = BN.host_inputs_detect(synth, 500_000, B, kind)
    pl, rl, dl = lay(p, (11,)), lay(ro, (4,)), lay(do, (4,))
    ws = net_tools.detect_workspace(table, B, BN.TOP_K, dev)
    for it in range(3):
        out = net_tools.decode_detected_bboxes(table, rl, dl, pl, workspace=ws, **kw)
    torch.cuda.synchronize()
    print(kind, "detections:", sum(int((out[0][c] != 0).sum()) for c in out[0]),
          "fallback segments:", int(net_tools.detect_fallback_flags(ws)[1:].sum()))
