"""Fixture for the losses (f-3): refine_loss and det_clf_loss of the UNMODIFIED reference
(utils/net_tools.py:492-623) executed over oracle/tf_shim on targets produced by the reference's own
refine_groundtruth / det_groundtruth.  The shim evaluates TF's float32 ops with NumPy (its reductions have
NumPy's order, exp / log are NumPy's), so the values are pinned to ~1e-6 relative, inside the 1e-5 tolerance
the tests use.  TEST INFRASTRUCTURE ONLY.   python -m oracle.gen_golden_loss"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                    # noqa: E402
from oracle.tf_shim import to_numpy               # noqa: E402


def main():
    ref = ref_loader.load_reference()
    tf, nt, cfg = ref.tf, ref.net_tools, ref.config
    img, feats = (128, 128), [(16, 16), (8, 8), (4, 4), (2, 2), (1, 1), (1, 1)]
    anchors = ref_loader.reference_anchors(ref, img, feats)
    rng = np.random.default_rng(41)
    B, G = 4, 9
    shapes = [(fh, fw, a[2].shape[0]) for (fh, fw), a in zip(feats, anchors)]
    outs = {k: [] for k in ("refine_gt", "refine_cb", "refine_lab", "refine_pos", "det_gt", "det_mask", "det_lab", "iou")}
    ro = [(rng.standard_normal((B,) + s + (4,)) * 0.1).astype(np.float32) for s in shapes]
    do = [(rng.standard_normal((B,) + s + (4,)) * 0.3).astype(np.float32) for s in shapes]
    clf = [(rng.standard_normal((B,) + s + (11,)) * 3).astype(np.float32) for s in shapes]
    for t in clf:
        t[..., 0] += 2.0
    per_image = []
    for b in range(B):
        c = rng.uniform(0.15, 0.85, size=(G, 2)); hw = np.exp(rng.uniform(np.log(0.08), np.log(0.6), size=(G, 2)))
        corner = np.clip(np.concatenate([c - hw / 2, c + hw / 2], 1), 0, 1).astype(np.float32)
        center = to_numpy(ref.common_tools.cornerBboxes_2_centerBboxes(tf.constant(corner)))
        labels = rng.integers(1, 11, size=G).astype(np.int64)
        r = nt.refine_groundtruth(anchors, tf.constant(center), tf.constant(labels, dtype=np.int64), cfg.refine_method.JACCARD_BIGGER)
        per_image.append([[to_numpy(t) for t in part] for part in r])
    for k, idx in (("refine_gt", 0), ("refine_cb", 1), ("refine_lab", 2), ("refine_pos", 3)):
        outs[k] = [np.stack([per_image[b][idx][l] for b in range(B)]) for l in range(len(shapes))]
    # make the ARM head good on the positives so that ODM positives exist
    for l in range(len(shapes)):
        pos = outs["refine_pos"][l].astype(bool)[..., 0]
        ro[l][pos] = outs["refine_gt"][l][pos] + (rng.standard_normal(outs["refine_gt"][l][pos].shape) * 0.05).astype(np.float32)
    T = lambda lst, dt=None: [tf.constant(a, dtype=dt) for a in lst]
    d = nt.det_groundtruth(T(ro), T(outs["refine_gt"]), T(outs["refine_cb"]), T(outs["refine_lab"], np.int32),
                           T(outs["refine_pos"], np.int32), anchors)
    for k, part in zip(("det_gt", "det_mask", "det_lab", "iou"), d):
        outs[k] = [to_numpy(t) for t in part]
    rl = nt.refine_loss(T(ro), T(outs["refine_gt"]), T(outs["refine_pos"], np.int32))
    dl, cl = nt.det_clf_loss(T(ro), T(clf), T(do), T(outs["det_gt"]), T(outs["det_mask"], np.int32), T(outs["det_lab"], np.int32),
                             T(outs["iou"]))
    z = {"n_layers": np.int64(len(shapes)), "refine_loss": np.float32(to_numpy(rl)), "det_loss": np.float32(to_numpy(dl)),
         "clf_loss": np.float32(to_numpy(cl))}
    for l in range(len(shapes)):
        z["ro_%d" % l], z["do_%d" % l], z["clf_%d" % l] = ro[l], do[l], clf[l]
        for k in outs:
            z["%s_%d" % (k, l)] = outs[k][l]
    print("refine_loss %.6f det_loss %.6f clf_loss %.6f  ARM pos %d ODM pos %d" % (
        z["refine_loss"], z["det_loss"], z["clf_loss"], sum(int(m.sum()) for m in outs["refine_pos"]), sum(int(m.sum()) for m in outs["det_mask"])))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "losses.npz"), **z)


if __name__ == "__main__":
    main()
