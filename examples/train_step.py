#!/usr/bin/env python
"""The box-level part of one training step, call for call as in the reference's train.py:95-152, on synthetic data:

    anchors_all = net_tools.anchors_all_layer(...)                                   train.py:96-99
    center_bboxes = cornerBboxes_2_centerBboxes(bboxes)                              :109
    refine_gt, refine_cbboxes, refine_labels, refine_pos_mask = refine_groundtruth() :111-113  (per image, then batched)
    refine_out, det_out, clf_out = net.get_output()                                  :139
    refine_loss(refine_out, refine_gt, refine_pos_mask)                              :142
    det_gt, det_pos_mask, det_labels, iou = det_groundtruth(...)                     :145-147
    det_loss, clf_loss = det_clf_loss(...)                                           :149-150
and the same targets from the fused extension `net_tools.target_gen` (one kernel)."""
import torch

from _synthetic import anchors_for, fake_network, ground_truth
from rodet_b200 import config
from rodet_b200.utils import net_tools, tf_extended as tfe                    # noqa: F401  (same imports as train.py:10-16)
from rodet_b200.utils.common_tools import cornerBboxes_2_centerBboxes
from rodet_b200.utils.data_pileline_tools import process_raw_gt_train


def run(device="cuda:0", batch=4, first_image=0):
    device = torch.device(device)
    anchors_all = anchors_for("mobilenet_v2")
    method = config.refine_method.JACCARD_BIGGER
    bboxes, labels, counts = ground_truth(first_image, batch, device)
    # GT-box half of prepare_data_train (crop -> overlap filter -> flip -> clamp), here with the identity crop
    labels, bboxes, counts = process_raw_gt_train(labels, bboxes, counts)
    center_bboxes = cornerBboxes_2_centerBboxes(bboxes)

    # the reference matches image by image and lets tf.train.batch stack the results (train.py:111-124)
    per_image = [net_tools.refine_groundtruth(anchors_all, center_bboxes[b, :int(counts[b])], labels[b, :int(counts[b])], method)
                 for b in range(batch)]
    stacked = [[torch.stack([per_image[b][k][l] for b in range(batch)]) for l in range(len(anchors_all))] for k in range(4)]
    # batched extension: one launch for the whole batch
    refine_gt, refine_cbboxes, refine_labels, refine_pos_mask = net_tools.refine_groundtruth(
        anchors_all, center_bboxes, labels, method, gt_counts=counts)
    same = all(torch.equal(a, b) for k, lst in enumerate((refine_gt, refine_cbboxes, refine_labels, refine_pos_mask))
               for a, b in zip(lst, stacked[k]))

    net = fake_network(batch, device, seed=first_image, refine_targets=(refine_gt, refine_pos_mask))
    refine_out, det_out, clf_out = net.get_output()
    leaves = [[t.clone().requires_grad_(True) for t in ts] for ts in (refine_out, det_out, clf_out)]
    refine_out, det_out, clf_out = leaves

    refine_loss = net_tools.refine_loss(refine_out, refine_gt, refine_pos_mask)
    det_gt, det_pos_mask, det_labels, iou_all_layers = net_tools.det_groundtruth(
        refine_out, refine_gt, refine_cbboxes, refine_labels, refine_pos_mask, anchors_all)
    det_loss, clf_loss = net_tools.det_clf_loss(refine_out, clf_out, det_out, det_gt, det_pos_mask, det_labels, iou_all_layers)
    total = refine_loss + det_loss + clf_loss
    total.backward()

    # fused extension: ARM + ODM targets in one kernel, bit-identical lists
    arm, det = net_tools.target_gen(anchors_all, center_bboxes, labels, [t.detach() for t in refine_out], gt_counts=counts)
    fused_same = all(torch.equal(torch.cat([x.reshape(batch, -1) for x in a], 1), torch.cat([x.reshape(batch, -1) for x in b], 1))
                     for a, b in zip(arm + det, (refine_gt, refine_cbboxes, refine_labels, refine_pos_mask,
                                                 det_gt, det_pos_mask, det_labels, iou_all_layers)))
    return {"per_image_equals_batched": bool(same), "fused_equals_two_calls": bool(fused_same),
            "arm_positives": int(sum(int(m.sum()) for m in refine_pos_mask)), "odm_positives": int(sum(int(m.sum()) for m in det_pos_mask)),
            "refine_loss": float(refine_loss.detach()), "det_loss": float(det_loss.detach()), "clf_loss": float(clf_loss.detach()),
            "grad_norm_refine": float(sum(t.grad.norm() ** 2 for t in refine_out) ** 0.5)}


if __name__ == "__main__":
    print(run())
