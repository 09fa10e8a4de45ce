"""Fixture for the eval TP/FP matching (f-2), produced by the UNMODIFIED reference
(`tfe.bboxes_matching_batch`, utils/tf_extended/bboxes.py:246-380) over oracle/tf_shim.
TEST INFRASTRUCTURE ONLY.   python -m oracle.gen_golden_eval"""
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
    rng = np.random.default_rng(91)
    B, N, G, classes = 3, 24, 7, [1, 2, 3]
    gb = np.sort(rng.uniform(0, 1, size=(B, G, 2, 2)).astype(np.float32), axis=2).reshape(B, G, 4)
    gl = rng.integers(1, 4, size=(B, G)).astype(np.int64)
    gl[:, -2:] = 0                                   # zero padding (dynamic_pad batches, evaluate.py:110-115)
    gb[:, -2:] = 0
    gd = (rng.uniform(size=(B, G)) < 0.2).astype(np.int64)
    gb[1, 1] = gb[1, 0]; gl[1, 1] = gl[1, 0]          # duplicate GT: argmax must take the first
    scores, bboxes = {}, {}
    for c in classes:
        s = np.sort(rng.uniform(0.2, 1, size=(B, N)).astype(np.float32), axis=1)[:, ::-1].copy()
        bx = np.zeros((B, N, 4), np.float32)
        for b in range(B):
            for i in range(N):
                own = np.nonzero(gl[b, :G - 2] == c)[0]
                g = int(rng.choice(own)) if len(own) and i % 4 else int(rng.integers(0, G - 2))
                jit = rng.normal(0, 0.012, size=4).astype(np.float32) if i % 3 else 0
                bx[b, i] = gb[b, g] + jit               # near (or exactly on) a GT box; repeats => double matches
        s[:, -5:] = 0; bx[:, -5:] = 0                   # pad_axis zero padding of the NMS output
        scores[c], bboxes[c] = s, bx
    n, tp, fp, sc = tfe.bboxes_matching_batch(classes, {c: tf.constant(v) for c, v in scores.items()},
                                              {c: tf.constant(v) for c, v in bboxes.items()},
                                              tf.constant(gl, dtype=np.int64), tf.constant(gb),
                                              tf.constant(gd, dtype=np.int64), matching_threshold=0.5)
    out = {"glabels": gl, "gbboxes": gb, "gdifficults": gd, "classes": np.asarray(classes), "thr": np.float32(0.5)}
    for c in classes:
        out["scores_c%d" % c] = scores[c]; out["bboxes_c%d" % c] = bboxes[c]
        out["n_c%d" % c] = to_numpy(n[c]); out["tp_c%d" % c] = to_numpy(tp[c]); out["fp_c%d" % c] = to_numpy(fp[c])
        print(c, out["n_c%d" % c], out["tp_c%d" % c].sum(), out["fp_c%d" % c].sum())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_matching.npz"), **out)


if __name__ == "__main__":
    main()
