"""Fixture for the evaluation metrics (f-2): tfe.precision_recall, average_precision_voc07 / voc12 and
cummax of the UNMODIFIED reference (utils/tf_extended/metrics.py:100-130, 210-258; math.py:41-67)
executed over oracle/tf_shim.  streaming_tp_fp_arrays itself needs TF variables and is restated
(oracle/restated.StreamingTpFp).  TEST INFRASTRUCTURE ONLY.   python -m oracle.gen_golden_metrics"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader                    # noqa: E402
from oracle.tf_shim import to_numpy               # noqa: E402


def cases():
    rng = np.random.default_rng(17)
    out = []
    for n, ngb, tie in ((0, 5, False), (1, 1, False), (37, 12, False), (300, 90, True), (5000, 1400, True), (64, 0, False)):
        s = rng.uniform(0.05, 1, size=n).astype(np.float32)
        if tie and n:
            s = (np.round(s * 50) / 50).astype(np.float32)          # many exactly equal scores: order decides
        tp = rng.uniform(size=n) < (0.25 + 0.6 * s)                  # better scores are right more often
        fp = ~tp
        drop = rng.uniform(size=n) < 0.1                             # some detections are neither (difficult GT)
        tp &= ~drop
        fp &= ~drop
        out.append((s, tp, fp, ngb))
    return out


def main():
    ref = ref_loader.load_reference()
    tf, tfe = ref.tf, ref.tfe
    z = {}
    for i, (s, tp, fp, ngb) in enumerate(cases()):
        p, r = tfe.precision_recall(tf.constant(np.int64(ngb)), tf.constant(np.int32(s.size)), tf.constant(tp), tf.constant(fp),
                                    tf.constant(s))
        # the float64 tensors go straight back in (tf.constant would apply TF's float32 default)
        z["voc07_%d" % i] = np.float64(to_numpy(tfe.average_precision_voc07(p, r)))
        z["voc12_%d" % i] = np.float64(to_numpy(tfe.average_precision_voc12(p, r)))
        p, r = to_numpy(p), to_numpy(r)
        assert p.dtype == np.float64
        z["scores_%d" % i], z["tp_%d" % i], z["fp_%d" % i], z["ngb_%d" % i] = s, tp, fp, np.int64(ngb)
        z["precision_%d" % i], z["recall_%d" % i] = p, r
        print(i, s.size, ngb, z["voc07_%d" % i], z["voc12_%d" % i])
    x = np.asarray([0.2, 0.9, 0.4, 0.95, 0.1, 0.95, 0.3])
    xt = tf.constant(x, dtype=np.float64)
    z["cummax_in"], z["cummax_fwd"], z["cummax_rev"] = x, to_numpy(tfe.cummax(xt)), to_numpy(tfe.cummax(xt, reverse=True))
    z["n_cases"] = np.int64(len(cases()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eval_metrics.npz"), **z)


if __name__ == "__main__":
    main()
