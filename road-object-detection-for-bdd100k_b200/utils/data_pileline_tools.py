"""The ground-truth-box half of the reference's training input pipeline
(`utils/data_pileline_tools.py:73-109`, file name as upstream).  Image decoding, cropping, resizing
and colour distortion stay outside the box-level path; what reaches
`refine_groundtruth` is the box chain

    bboxes = tfe.bboxes_resize(distort_bbox, bboxes)                       process.py:135
    labels, bboxes = tfe.bboxes_filter_overlap(labels, bboxes, 0.3, False) process.py:136-138
    bboxes = flip_bboxes(bboxes) if the image was mirrored                  tf_image.py:284-306
    bboxes = tf.minimum(tf.maximum(bboxes, 0.), 1.)                         data_pileline_tools.py:107-108

which `process_raw_gt_train` runs for a whole padded batch in one kernel (`rod_gt_boxes_update`)."""
from __future__ import annotations

from .tf_extended.bboxes import _gt_update

BBOX_CROP_OVERLAP = 0.3     # utils/augmentation/process.py:6


def process_raw_gt_train(labels, bboxes, counts=None, distort_bbox=None, mirror=None,
                         crop_overlap=BBOX_CROP_OVERLAP, assign_negative=False):
    """labels [B,G] int64/int32, bboxes [B,G,4] corner form (zero padded), counts [B] valid boxes per image
    (None: all G), distort_bbox [B,4] the crop sampled for each image (None: no crop), mirror [B] bool
    (None: no flip).  Returns (labels, bboxes, counts) after the chain above: kept boxes in order, rows
    zero padded, new counts int32 — ready for cornerBboxes_2_centerBboxes + refine_groundtruth(gt_counts=...).
    The overlap filter belongs to the crop: the reference applies it exactly when a crop was sampled, always at
    BBOX_CROP_OVERLAP (process.py:134-138).  So without `distort_bbox` nothing is filtered, and a crop with
    `crop_overlap=None` — a combination the reference does not have — is rejected instead of being guessed."""
    if distort_bbox is not None and crop_overlap is None:
        raise ValueError("crop_overlap=None with a distort_bbox: the reference always filters a crop's boxes "
                         "(utils/augmentation/process.py:136-138, threshold %.1f)" % BBOX_CROP_OVERLAP)
    return _gt_update(labels, bboxes, counts, distort_bbox, mirror, distort_bbox is not None,
                      0.0 if crop_overlap is None else crop_overlap, assign_negative, True)
