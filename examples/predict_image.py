#!/usr/bin/env python
"""The box-level part of the reference's predict.py:67-137 for ONE image on synthetic data:
ARM matching of the ground truth for display (predict.py:81-85), decode + post-process with the demo
thresholds 0.1 / 0.4 / 400 / 200 (:127-137)."""
import torch

from _synthetic import anchors_for, fake_network, ground_truth
from rodet_b200 import config
from rodet_b200.utils import net_tools
from rodet_b200.utils.common_tools import centerBboxes_2_cornerBboxes, cornerBboxes_2_centerBboxes


def run(device="cuda:0", image=7):
    device = torch.device(device)
    anchors_all = anchors_for("mobilenet_v2")
    bboxes, labels, counts = ground_truth(image, 1, device)
    g = int(counts[0])
    center_bboxes = cornerBboxes_2_centerBboxes(bboxes[0, :g])
    refine_gt, refine_cbboxes, refine_labels, refine_pos_mask = net_tools.refine_groundtruth(
        anchors_all, center_bboxes, labels[0, :g], config.refine_method.JACCARD_BIGGER)
    # predict.py:93-105: matched GT boxes / labels of every anchor, flattened over the layers
    corner_bboxes_gt = torch.cat([centerBboxes_2_cornerBboxes(box).reshape(-1, 4) for box in refine_cbboxes], 0)
    labels_gt = torch.cat([lab.reshape(-1) for lab in refine_labels], 0)

    net = fake_network(1, device, seed=image)
    refine_out, det_out, clf_out = net.get_output()
    predition_all_layers = [net_tools.softmax(clf) for clf in clf_out]
    locations_all_layers = [centerBboxes_2_cornerBboxes(net_tools.decode_locations_one_layer(a, ro + do))
                            for ro, do, a in zip(refine_out, det_out, anchors_all)]
    rscores, rbboxes = net_tools.detected_bboxes(predition_all_layers, locations_all_layers, select_threshold=0.1,
                                                 nms_threshold=0.4, top_k=400, keep_top_k=200)
    return {"gt_boxes": g, "matched_anchors": int((labels_gt > 0).sum()), "matched_box_rows": int(corner_bboxes_gt.shape[0]),
            "detections_per_class": {int(c): int((rscores[c] > 0).sum()) for c in rscores}}


if __name__ == "__main__":
    print(run())
