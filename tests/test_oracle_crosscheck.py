"""Independent cross-checks of the oracle's restatement of the TensorFlow kernels whose source is not in the
reference tree (tf.nn.top_k, tf.image.non_max_suppression — "parity unpinned" there):

  * oracle.restated.nms_indices against torchvision.ops.nms (CPU) on valid boxes with distinct scores — the
    secondary cross-check SURVEY 8c names (torchvision differs on zero-area boxes and ties, which are left out);
  * against a literal pure-Python greedy loop written from the adopted semantics (descending score, ties ->
    lowest index, suppress iff IoU > thr, IoU 0 when an area <= 0, stop at keep_top_k) on small cases WITH
    duplicate scores, duplicate boxes and zero-area boxes;
  * oracle.restated.topk_indices against Python's sorted() on (-score, index).

CPU only; nothing here touches the product package."""
import numpy as np
import pytest

from oracle import restated as R

f32 = np.float32


def _boxes(rng, n, valid=True):
    c = rng.uniform(0.1, 0.9, size=(n, 2))
    hw = np.exp(rng.uniform(np.log(0.03), np.log(0.5), size=(n, 2)))
    b = np.concatenate([c - hw / 2, c + hw / 2], 1).astype(f32)        # ymin, xmin, ymax, xmax
    if not valid:
        k = max(1, n // 8)
        idx = rng.choice(n, size=k, replace=False)
        b[idx, 2] = b[idx, 0]                                          # zero height
        swap = rng.choice(n, size=k, replace=False)
        b[swap] = b[swap][:, [2, 1, 0, 3]]                             # ymin / ymax swapped: normalised by min / max
    return b


def _iou_literal(a, b):
    """TF NonMaxSuppression IOU in float32 scalars."""
    ay0, ay1 = min(a[0], a[2]), max(a[0], a[2]); ax0, ax1 = min(a[1], a[3]), max(a[1], a[3])
    by0, by1 = min(b[0], b[2]), max(b[0], b[2]); bx0, bx1 = min(b[1], b[3]), max(b[1], b[3])
    area_a = f32(f32(ay1 - ay0) * f32(ax1 - ax0)); area_b = f32(f32(by1 - by0) * f32(bx1 - bx0))
    if area_a <= 0 or area_b <= 0:
        return f32(0)
    ih = max(f32(min(ay1, by1) - max(ay0, by0)), f32(0)); iw = max(f32(min(ax1, bx1) - max(ax0, bx0)), f32(0))
    inter = f32(ih * iw)
    return f32(inter / f32(f32(area_a + area_b) - inter))


def _nms_literal(scores, boxes, thr, keep):
    order = sorted(range(len(scores)), key=lambda i: (-float(scores[i]), i))
    sel = []
    for i in order:
        if len(sel) >= keep:
            break
        if all(not (_iou_literal(boxes[j], boxes[i]) > f32(thr)) for j in sel):
            sel.append(i)
    return np.asarray(sel, dtype=np.int32)


@pytest.mark.parametrize("n,thr", [(50, 0.3), (200, 0.45), (400, 0.45), (400, 0.6)])
def test_nms_matches_torchvision_on_valid_boxes(n, thr):
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    for trial in range(8):
        rng = np.random.default_rng([n, int(thr * 100), trial])
        boxes = _boxes(rng, n)
        scores = rng.permutation(n).astype(f32) / f32(n) + f32(0.001)          # distinct
        got = R.nms_indices(scores, boxes, thr, n)
        xyxy = torch.from_numpy(boxes[:, [1, 0, 3, 2]].copy())
        want = tv.ops.nms(xyxy, torch.from_numpy(scores), float(thr)).numpy().astype(np.int32)
        assert np.array_equal(got, want), (n, thr, trial)
        if trial < 2:
            # Tier A's kernel (oracle/tf_shim, the literal loop behind the golden fixtures) as well
            from oracle import tf_shim
            tf = tf_shim.install()
            shim = tf_shim.to_numpy(tf.image.non_max_suppression(boxes, scores, n, thr))
            assert np.array_equal(shim, want), ("tf_shim", n, thr, trial)


@pytest.mark.parametrize("seed", range(12))
def test_nms_matches_literal_loop_with_ties_and_degenerate_boxes(seed):
    rng = np.random.default_rng([77, seed])
    n = int(rng.integers(5, 60))
    boxes = _boxes(rng, n, valid=False)
    dup = rng.choice(n, size=max(1, n // 5), replace=False)
    boxes[dup] = boxes[rng.choice(n, size=len(dup))]                            # duplicate boxes (IoU exactly 1)
    scores = (rng.integers(0, 6, size=n).astype(f32) / f32(5))                  # heavy score ties, zeros included
    for thr in (0.0, 0.45, 1.0):
        for keep in (3, n):
            got = R.nms_indices(scores, boxes, thr, keep)
            assert np.array_equal(got, _nms_literal(scores, boxes, thr, keep)), (seed, thr, keep)
    # and the padded form the reference returns (tensors.py:59-86): selection order, zeros behind it
    s, b = R.bboxes_nms(scores, boxes, 0.45, n)
    k = len(_nms_literal(scores, boxes, 0.45, n))
    assert s.shape == (n,) and np.all(s[k:] == 0) and np.all(b[k:] == 0)


@pytest.mark.parametrize("seed", range(6))
def test_topk_ties_take_the_lower_index(seed):
    rng = np.random.default_rng([5, seed])
    n = int(rng.integers(10, 300))
    s = (rng.integers(0, 8, size=(3, n)).astype(f32) / f32(7))
    for k in (1, n // 2, n):
        got = R.topk_indices(s, k)
        for r in range(3):
            want = sorted(range(n), key=lambda i: (-float(s[r, i]), i))[:k]
            assert list(got[r]) == want
