#!/usr/bin/env python
"""The box-level part of the reference's evaluate.py:128-208 on synthetic data, call for call:

    predition_all_layers[i] = softmax(clf_out[i])                                    evaluate.py:136-137
    corner = centerBboxes_2_cornerBboxes(decode_locations_one_layer(anchors_l, refine_out_l + det_out_l))   :139-143
    rscores, rbboxes = net_tools.detected_bboxes(..., select_threshold, nms_threshold, top_k, keep_top_k)   :147-151
    num_gbboxes, tp, fp, rscores = tfe.bboxes_matching_batch(...)                    :154-157
    tfe.streaming_tp_fp_arrays -> tfe.precision_recall -> tfe.average_precision_voc07 / voc12 -> mAP       :168-208
plus the one-call extension `decode_detected_bboxes(..., from_logits=True)` (softmax and decode fused in)."""
import torch

from _synthetic import anchors_for, fake_network, ground_truth
from rodet_b200.utils import net_tools, tf_extended as tfe
from rodet_b200.utils.common_tools import centerBboxes_2_cornerBboxes

SELECT_THRESHOLD, NMS_THRESHOLD, SELECT_TOP_K, KEEP_TOP_K, MATCHING_THRESHOLD = 0.3, 0.4, 400, 200, 0.5   # evaluate.py:58-67


def run(device="cuda:0", batch=4, num_batches=2):
    device = torch.device(device)
    anchors_all = anchors_for("mobilenet_v2")
    tfe.reset_local_variables()
    fused_same = True
    for it in range(num_batches):
        b_gbboxes, b_glabels, counts = ground_truth(1000 + it * batch, batch, device)
        b_difficults = torch.zeros_like(b_glabels)
        net = fake_network(batch, device, seed=100 + it)
        refine_out, det_out, clf_out = net.get_output()

        predition_all_layers = [net_tools.softmax(clf) for clf in clf_out]
        locations_all_layers = []
        for ro, do, anchors_one_layer in zip(refine_out, det_out, anchors_all):
            center_locations = net_tools.decode_locations_one_layer(anchors_one_layer, ro + do)
            locations_all_layers.append(centerBboxes_2_cornerBboxes(center_locations))

        rscores, rbboxes = net_tools.detected_bboxes(predition_all_layers, locations_all_layers,
                                                     select_threshold=SELECT_THRESHOLD, nms_threshold=NMS_THRESHOLD,
                                                     top_k=SELECT_TOP_K, keep_top_k=KEEP_TOP_K)
        fs, fb = net_tools.decode_detected_bboxes(anchors_all, refine_out, det_out, clf_out, select_threshold=SELECT_THRESHOLD,
                                                  nms_threshold=NMS_THRESHOLD, top_k=SELECT_TOP_K, keep_top_k=KEEP_TOP_K,
                                                  from_logits=True)
        fused_same &= all(torch.equal(fs[c], rscores[c]) and torch.equal(fb[c], rbboxes[c]) for c in rscores)

        num_gbboxes, tp, fp, rscores = tfe.bboxes_matching_batch(rscores.keys(), rscores, rbboxes, b_glabels, b_gbboxes,
                                                                  b_difficults, matching_threshold=MATCHING_THRESHOLD)
        tp_fp_metric = tfe.streaming_tp_fp_arrays(num_gbboxes, tp, fp, rscores)

    aps_voc07, aps_voc12 = {}, {}
    for c in tp_fp_metric[0].keys():
        prec, rec = tfe.precision_recall(*tp_fp_metric[0][c])
        aps_voc07[c] = float(tfe.average_precision_voc07(prec, rec))
        aps_voc12[c] = float(tfe.average_precision_voc12(prec, rec))
    return {"mAP_VOC07": sum(aps_voc07.values()) / len(aps_voc07), "mAP_VOC12": sum(aps_voc12.values()) / len(aps_voc12),
            "detections": int(sum(int(tp_fp_metric[0][c][1]) for c in tp_fp_metric[0])), "fused_equals_call_sequence": bool(fused_same)}


if __name__ == "__main__":
    print(run())
