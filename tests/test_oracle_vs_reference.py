"""Live cross-check (build container only): the UNMODIFIED reference sources executed over
oracle/tf_shim against oracle/restated.py on fresh random inputs.  Skipped where /root/reference
does not exist (the GPU box); the committed fixtures in tests/golden cover that case."""
import numpy as np
import pytest

from oracle import ref_loader
from oracle import restated as R

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference sources not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load_reference()


def _flat(per_layer, tail):
    outs = []
    for t in per_layer:
        t = np.asarray(t)
        outs.append(t.reshape((-1,) + t.shape[t.ndim - tail:]) if tail else t.reshape(-1))
    return np.concatenate(outs, axis=0)


@pytest.mark.parametrize("seed,g", [(101, 4), (102, 23)])
def test_arm_live(ref, seed, g):
    from oracle.tf_shim import to_numpy
    img, feats = (160, 160), [(20, 20), (10, 10), (5, 5), (3, 3), (2, 2), (1, 1)]
    anc = ref_loader.reference_anchors(ref, img, feats)
    table = R.AnchorTable(anc)
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.1, 0.9, size=(g, 2))
    hw = np.exp(rng.uniform(np.log(0.04), np.log(0.6), size=(g, 2)))
    corner = np.clip(np.concatenate([c - hw / 2, c + hw / 2], 1), 0, 1).astype(np.float32)
    labels = rng.integers(1, 11, size=g).astype(np.int64)
    tf = ref.tf
    center = to_numpy(ref.common_tools.cornerBboxes_2_centerBboxes(tf.constant(corner)))
    out = ref.net_tools.refine_groundtruth(anc, tf.constant(center), tf.constant(labels, dtype=np.int64),
                                           ref.config.refine_method.JACCARD_BIGGER)
    gt, cb, lab, pos = [to_numpy(v) for v in out]
    o = R.arm_match_encode(table, center, labels)
    assert np.array_equal(_flat(pos, 1)[:, 0], o[3]) and np.array_equal(_flat(lab, 1)[:, 0], o[2])
    assert np.array_equal(_flat(cb, 1), o[1])
    np.testing.assert_allclose(_flat(gt, 1), o[0], rtol=1e-5, atol=1e-6)
    assert o[3].sum() > 0


def test_detect_live(ref):
    from oracle.tf_shim import to_numpy
    img, feats = (96, 96), [(6, 5), (3, 3), (2, 2), (1, 2), (1, 1), (1, 1)]
    anc = ref_loader.reference_anchors(ref, img, feats)
    table = R.AnchorTable(anc)
    rng = np.random.default_rng(7)
    B, n = 2, table.n
    probs = rng.uniform(0, 1, size=(B, n, 11)).astype(np.float32)
    probs = (np.round(probs * 16) / 16).astype(np.float32)
    boxes = np.sort(rng.uniform(0, 1, size=(B, n, 2, 2)).astype(np.float32), axis=2).reshape(B, n, 4)
    tf = ref.tf
    split = lambda a, tail: [tf.constant(np.ascontiguousarray(a[:, table.offsets[l]:table.offsets[l + 1]]).reshape(
        (B,) + table.shapes[l] + tail)) for l in range(6)]
    rs, rb = ref.net_tools.detected_bboxes(split(probs, (11,)), split(boxes, (4,)), select_threshold=0.5,
                                           nms_threshold=0.3, top_k=30, keep_top_k=12)
    o_s, o_b = R.detected_bboxes(probs, boxes, 0.5, 0.3, None, 30, 12)
    for c in range(1, 11):
        assert np.array_equal(to_numpy(rs[c]), o_s[c]) and np.array_equal(to_numpy(rb[c]), o_b[c])
