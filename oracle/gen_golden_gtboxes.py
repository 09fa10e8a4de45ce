"""Fixture for the GT-box chain of the training input pipeline (f-4): tfe.bboxes_resize and
tfe.bboxes_filter_overlap of the UNMODIFIED reference (utils/tf_extended/bboxes.py:139-163, 408-428) over
oracle/tf_shim, followed by the flip and clamp lines of tf_image.py:286-288 / data_pileline_tools.py:107-108
written with the same tf ops (flip_bboxes is a closure inside random_flip_left_right and cannot be called
on its own).  TEST INFRASTRUCTURE ONLY.   python -m oracle.gen_golden_gtboxes"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                    # noqa: E402
from oracle.tf_shim import to_numpy               # noqa: E402


def main():
    ref = ref_loader.load_reference()
    tf, tfe = ref.tf, ref.tfe
    rng = np.random.default_rng(29)
    B, G = 6, 40
    z = {"B": np.int64(B), "G": np.int64(G)}
    boxes = np.zeros((B, G, 4), np.float32)
    labels = np.zeros((B, G), np.int64)
    counts = np.asarray([40, 17, 1, 0, 33, 40], np.int32)
    crops = np.zeros((B, 4), np.float32)
    mirror = np.asarray([0, 1, 1, 0, 1, 0], bool)
    for b in range(B):
        c = rng.uniform(0.05, 0.95, size=(G, 2)); hw = np.exp(rng.uniform(np.log(0.02), np.log(0.6), size=(G, 2)))
        bx = np.clip(np.concatenate([c - hw / 2, c + hw / 2], 1), 0, 1).astype(np.float32)
        boxes[b, :counts[b]] = bx[:counts[b]]
        labels[b, :counts[b]] = rng.integers(1, 11, size=counts[b])
        y0, x0 = rng.uniform(0, 0.3, 2); y1, x1 = rng.uniform(0.6, 1.0, 2)
        crops[b] = [y0, x0, y1, x1]
    crops[5] = [0, 0, 1, 1]                                         # identity crop: nothing moves, boundary boxes stay
    for b in range(B):
        n = counts[b]
        lb, bx = tf.constant(labels[b, :n], dtype=np.int64), tf.constant(boxes[b, :n].reshape(n, 4))
        bx = tfe.bboxes_resize(tf.constant(crops[b]), bx)            # process.py:135
        for neg in (False, True):
            l2, b2 = tfe.bboxes_filter_overlap(lb, bx, threshold=0.3, assign_negative=neg)   # process.py:136-138
            if mirror[b]:
                b2 = tf.stack([b2[:, 0], 1 - b2[:, 3], b2[:, 2], 1 - b2[:, 1]], axis=-1)      # tf_image.py:286-288
            b2 = tf.minimum(tf.maximum(b2, 0.), 1.)                                           # data_pileline_tools.py:107-108
            z["labels_%d_%d" % (b, neg)] = to_numpy(l2).astype(np.int64).reshape(-1)
            z["bboxes_%d_%d" % (b, neg)] = to_numpy(b2).astype(np.float32).reshape(-1, 4)
        print(b, n, z["labels_%d_0" % b].size, int((z["labels_%d_1" % b] < 0).sum()))
    z.update(boxes=boxes, labels=labels, counts=counts, crops=crops, mirror=mirror)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gt_boxes.npz"), **z)


if __name__ == "__main__":
    main()
