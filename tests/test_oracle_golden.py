"""Tier B (oracle/restated.py) against the fixtures produced by running the
UNMODIFIED reference over the TF shim (oracle/gen_golden.py).  CPU only.

Bar: integer outputs (pos masks, labels, argmax index), matched boxes and NMS keep
sets bit-exact; exp/log-dependent floats within 1e-5 relative (abs floor 1e-6)."""
import hashlib

import numpy as np
import pytest

from conftest import LAYOUTS, golden, golden_anchors, golden_files
from helpers import bit_equal
from oracle import restated as R

RTOL, ATOL = 1e-5, 1e-6


def _digest(a):
    a = np.ascontiguousarray(a)
    h = hashlib.sha256()
    h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("layout", ["418", "512", "tiny"])
def test_anchor_generation_bit_exact(layout):
    img, feats = LAYOUTS[layout]
    ours = R.anchors_all_layer(img, feats)
    ref = golden_anchors(layout)
    for (y, x, h, w), (ry, rx, rh, rw) in zip(ours, ref):
        for a, b in ((y, ry), (x, rx), (h, rh), (w, rw)):
            assert a.dtype == np.float32 and a.shape == b.shape
            assert np.array_equal(a, b)
    z = golden("anchors.npz")
    sizes = np.concatenate(list(R.init_anchor(6, img).values()))
    assert np.array_equal(sizes, z["%s_sizes_px" % layout])
    assert list(z["n_anchor_each_layer"]) == [6, 9, 9, 9, 9, 9]


def test_anchor_counts():
    assert R.AnchorTable(golden_anchors("418")).n == 25800
    assert R.AnchorTable(golden_anchors("512")).n == 36852


def test_box_format():
    z = golden("box_format.npz")
    assert np.array_equal(R.corner_to_center(z["corner"]), z["center"])
    assert np.array_equal(R.center_to_corner(z["center"]), z["corner_back"])


@pytest.mark.parametrize("fname", golden_files("targets_"))
def test_targets(fname):
    z = golden(fname)
    layout = str(z["layout"])
    table = R.AnchorTable(golden_anchors(layout))
    center, labels = z["center"], z["labels"]
    assert np.array_equal(R.corner_to_center(z["corner"]), center)
    gt, cb, lab, pos, idx = R.arm_match_encode(table, center, labels)
    assert np.array_equal(pos, z["jb_pos"])
    assert np.array_equal(lab, z["jb_labels"])
    assert np.array_equal(idx, z["jb_idx"])
    if "jb_gt" in z:
        assert np.array_equal(cb, z["jb_cb"])
        np.testing.assert_allclose(gt, z["jb_gt"], rtol=RTOL, atol=ATOL)
        ref_gt, ref_cb = z["jb_gt"], z["jb_cb"]
    else:
        assert _digest(cb) == str(z["jb_cb_digest"])
        np.testing.assert_allclose(gt[pos > 0], z["jb_gt_sparse"], rtol=RTOL, atol=ATOL)
        assert not gt[pos == 0].any()
        ref_gt = np.zeros_like(gt); ref_gt[pos > 0] = z["jb_gt_sparse"]
        ref_cb = cb
    if "nn_gt" in z:
        ngt, ncb, nlab, npos, _ = R.arm_match_encode(table, center, labels, method="NEAREST_NEIGHBOR")
        assert np.array_equal(npos, z["nn_pos"]) and np.array_equal(nlab, z["nn_labels"])
        assert np.array_equal(ncb, z["nn_cb"])
        np.testing.assert_allclose(ngt, z["nn_gt"], rtol=RTOL, atol=ATOL)
    if "enc_gt0" in z:
        np.testing.assert_allclose(R.encode(table.center, center[0]), z["enc_gt0"], rtol=RTOL, atol=ATOL)

    # ODM, fed the REFERENCE's ARM outputs (stage-wise parity)
    B = z["odm_mask"].shape[0]
    if "odm_refine_out" in z:
        ro = z["odm_refine_out"]
    else:
        pytest.skip("full-size ODM/decode inputs are regenerated in test_targets_fullsize")
    og = np.broadcast_to(ref_gt, (B,) + ref_gt.shape)
    det_gt, m, dl, iou = R.odm_target(table, ro, og, np.broadcast_to(ref_cb, og.shape),
                                      np.broadcast_to(z["jb_labels"], (B, table.n)),
                                      np.broadcast_to(z["jb_pos"], (B, table.n)))
    assert np.array_equal(m, z["odm_mask"])
    assert np.array_equal(dl, z["odm_labels"])
    np.testing.assert_allclose(iou, z["odm_iou"], rtol=1e-4, atol=ATOL)
    assert np.array_equal(det_gt, z["odm_det_gt"])
    np.testing.assert_allclose(R.decode_corner(table, ro, z["dec_det_out"]), z["dec_corner"],
                               rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("fname", [f for f in golden_files("targets_r")])
def test_targets_fullsize(fname):
    """Full-size fixtures keep only seeds + digests of the random inputs: regenerate
    them exactly as oracle/gen_golden.py does and check the digests first."""
    z = golden(fname)
    layout = str(z["layout"])
    table = R.AnchorTable(golden_anchors(layout))
    rng = np.random.default_rng(int(z["seed"]))
    g = z["corner"].shape[0]
    rng.uniform(0.1, 0.9, size=(g, 2)); rng.uniform(np.log(0.03), np.log(0.6), size=(g, 2))
    rng.integers(1, 11, size=g)
    n, B = table.n, 2
    gt, cb, lab, pos, _ = R.arm_match_encode(table, z["center"], z["labels"])
    ref_gt = np.zeros_like(gt); ref_gt[pos > 0] = z["jb_gt_sparse"]
    ro = (rng.standard_normal(size=(B, n, 4)) * np.array([0.1, 0.1, 0.2, 0.2])).astype(np.float32)
    ro = np.where(pos[None, :, None] > 0,
                  (ref_gt[None] + rng.uniform(0, 2.5, size=(B, n, 1)) * ro).astype(np.float32), ro)
    assert _digest(ro) == str(z["odm_refine_out_digest"]), "fixture inputs could not be regenerated"
    og = np.broadcast_to(ref_gt, (B, n, 4))
    det_gt, m, dl, iou = R.odm_target(table, ro, og, np.broadcast_to(cb, og.shape),
                                      np.broadcast_to(lab, (B, n)), np.broadcast_to(pos, (B, n)))
    assert np.array_equal(m, z["odm_mask"])
    assert np.array_equal(dl, z["odm_labels"])
    np.testing.assert_allclose(iou, z["odm_iou"], rtol=1e-4, atol=ATOL)
    assert np.array_equal(det_gt[m > 0], z["odm_det_gt_sparse"])
    do = (rng.standard_normal(size=(B, n, 4)) * np.array([0.1, 0.1, 0.2, 0.2])).astype(np.float32)
    assert _digest(do) == str(z["dec_det_out_digest"])
    np.testing.assert_allclose(R.decode_corner(table, ro, do)[:, ::37], z["dec_corner"], rtol=RTOL, atol=ATOL)


def _regen_detect_inputs(z, table):
    rng = np.random.default_rng(int(z["seed"]))
    B, n = int(z["B"]), table.n
    mode = str(z["mode"])
    if mode == "stress":
        probs = rng.uniform(round(float(z["select_threshold"]), 6), 1.0, size=(B, n, 11)).astype(np.float32)
        sig = np.array([0.05] * 4)
    else:
        zz = (rng.standard_normal(size=(B, n, 11)) * 3.0).astype(np.float32)
        zz[..., 0] += np.float32(4.0)
        zz -= zz.max(-1, keepdims=True)
        e = np.exp(zz)
        probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
        sig = np.array([0.1, 0.1, 0.2, 0.2])
    ro = (rng.standard_normal(size=(B, n, 4)) * sig).astype(np.float32)
    do = (rng.standard_normal(size=(B, n, 4)) * sig).astype(np.float32)
    assert _digest(probs) == str(z["probs_digest"]), "fixture inputs could not be regenerated"
    assert _digest(ro) == str(z["refine_out_digest"]) and _digest(do) == str(z["det_out_digest"])
    return probs, ro, do


@pytest.mark.parametrize("fname", golden_files("detect_"))
def test_detected_bboxes(fname):
    z = golden(fname)
    layout = str(z["layout"])
    table = R.AnchorTable(golden_anchors(layout))
    if "probs" in z:
        probs, ro, do = z["probs"], z["refine_out"], z["det_out"]
    else:
        probs, ro, do = _regen_detect_inputs(z, table)
    boxes = R.decode_corner(table, ro, do)
    if "boxes_corner" in z:
        np.testing.assert_allclose(boxes, z["boxes_corner"], rtol=RTOL, atol=ATOL)
        boxes = z["boxes_corner"]           # stage-wise: feed the reference's boxes
    sthr = None if bool(z["select_none"]) else float(z["select_threshold"])
    clip = z["clip"] if "clip" in z else None
    rs, rb = R.detected_bboxes(probs, boxes, sthr, float(z["nms_threshold"]), clip,
                               int(z["top_k"]), int(z["keep_top_k"]))
    exact = "boxes_corner" in z
    for c in range(1, 11):
        if exact:
            assert np.array_equal(rs[c], z["scores_c%d" % c]), c
            assert np.array_equal(rb[c], z["bboxes_c%d" % c]), c
        else:   # boxes come from our own decode (exp within 1 ulp of the reference's)
            assert np.array_equal(rs[c], z["scores_c%d" % c]), c
            np.testing.assert_allclose(rb[c], z["bboxes_c%d" % c], rtol=RTOL, atol=ATOL)
    if "sel_scores_c1" in z:
        d_s, d_b = R.bboxes_select(probs, boxes, sthr)
        s_s, s_b = R.bboxes_sort(d_s, d_b, int(z["top_k"]))
        for c in range(1, 11):
            assert np.array_equal(d_s[c], z["sel_scores_c%d" % c])
            assert np.array_equal(d_b[c], z["sel_bboxes_c%d" % c])
            assert np.array_equal(s_s[c], z["sort_scores_c%d" % c])
            assert np.array_equal(s_b[c], z["sort_bboxes_c%d" % c])


def test_tfe_ops():
    z = golden("tfe_ops.npz")
    b, r, s = z["boxes"], z["ref"], z["scores"]
    assert np.array_equal(R.bboxes_jaccard(r, b), z["jaccard"])
    assert np.array_equal(R.bboxes_intersection(r, b), z["intersection"])
    assert np.array_equal(R.bboxes_resize(r, b), z["resize"])
    assert np.array_equal(R.bboxes_clip(r, b), z["clip"])
    ns, nb = R.bboxes_nms(s, b, 0.3, 25)
    assert np.array_equal(ns, z["nms_scores"]) and np.array_equal(nb, z["nms_bboxes"])
    ss, sb = R.bboxes_sort(s[None], b[None], 20)
    assert np.array_equal(ss, z["sort_scores"]) and np.array_equal(sb, z["sort_bboxes"])


def test_eval_matching_restatement():
    """f-2: tfe.bboxes_matching_batch restated, against the reference fixture."""
    z = golden("eval_matching.npz")
    cl = [int(c) for c in z["classes"]]
    n, tp, fp, _ = R.bboxes_matching_batch(cl, {c: z["scores_c%d" % c] for c in cl}, {c: z["bboxes_c%d" % c] for c in cl},
                                           z["glabels"], z["gbboxes"], z["gdifficults"], float(z["thr"]))
    for c in cl:
        assert np.array_equal(n[c], z["n_c%d" % c])
        assert np.array_equal(tp[c], z["tp_c%d" % c]) and np.array_equal(fp[c], z["fp_c%d" % c])
    assert sum(int(z["tp_c%d" % c].sum()) for c in cl) >= 8


def test_eval_metrics_restatement():
    """oracle/restated precision_recall / AP / cummax against the unmodified reference
    (utils/tf_extended/metrics.py:100-130, 210-258; math.py:41-67) run by oracle/gen_golden_metrics.py."""
    z = golden("eval_metrics.npz")
    for i in range(int(z["n_cases"])):
        s, tp, fp, ngb = z["scores_%d" % i], z["tp_%d" % i], z["fp_%d" % i], int(z["ngb_%d" % i])
        p, r = R.precision_recall(ngb, s.size, tp, fp, s)
        assert p.dtype == np.float64 and np.array_equal(p, z["precision_%d" % i]) and np.array_equal(r, z["recall_%d" % i])
        assert R.average_precision_voc07(p, r) == z["voc07_%d" % i]
        assert abs(R.average_precision_voc12(p, r) - z["voc12_%d" % i]) <= 1e-12 * max(1.0, abs(z["voc12_%d" % i]))
    assert np.array_equal(R.cummax(z["cummax_in"]), z["cummax_fwd"])
    assert np.array_equal(R.cummax(z["cummax_in"], reverse=True), z["cummax_rev"])
    # streaming accumulation: filter (tp | fp) & score > 1e-4 only when remove_zero_scores
    st = R.StreamingTpFp()
    sc = np.asarray([[0.9, 0.5, 0.0], [0.00005, 0.3, 0.2]], np.float32)
    tp = np.asarray([[1, 0, 0], [1, 0, 1]], bool)
    fp = np.asarray([[0, 1, 0], [0, 0, 0]], bool)
    st.update(np.asarray([2, 1]), tp, fp, sc)
    no, nd, t, f, s = st.value()
    assert int(no) == 3 and int(nd) == 3 and s.tolist() == [np.float32(0.9), np.float32(0.5), np.float32(0.2)]
    assert t.tolist() == [True, False, True] and f.tolist() == [False, True, False]
    st.update(np.asarray([4]), tp, fp, sc, remove_zero_scores=False)
    assert int(st.value()[0]) == 7 and int(st.value()[1]) == 9 and st.value()[4].size == 9


def test_gt_boxes_pipeline_restatement():
    """oracle/restated.gt_boxes_train against the unmodified tfe.bboxes_resize + bboxes_filter_overlap
    (+ the flip / clamp lines) run by oracle/gen_golden_gtboxes.py."""
    z = golden("gt_boxes.npz")
    for b in range(int(z["B"])):
        n = int(z["counts"][b])
        for neg in (0, 1):
            lab, bx = R.gt_boxes_train(z["labels"][b, :n], z["boxes"][b, :n], z["crops"][b], bool(z["mirror"][b]), 0.3, bool(neg))
            assert np.array_equal(lab, z["labels_%d_%d" % (b, neg)]), (b, neg)
            assert bit_equal(bx.reshape(-1, 4), z["bboxes_%d_%d" % (b, neg)]), (b, neg)


def test_losses_restatement():
    """oracle/restated smooth_l1_loss / clf_loss against refine_loss / det_clf_loss of the unmodified reference
    (utils/net_tools.py:492-623) run by oracle/gen_golden_loss.py (float tolerance 1e-5)."""
    z = golden("losses.npz")
    L = int(z["n_layers"])
    get = lambda k: [z["%s_%d" % (k, l)] for l in range(L)]
    rl = R.smooth_l1_loss(get("refine_gt"), get("ro"), get("refine_pos"))
    dl = R.smooth_l1_loss(get("det_gt"), get("do"), get("det_mask"))
    cl = R.clf_loss(get("clf"), get("det_mask"), get("det_lab"), get("iou"))
    assert abs(rl - float(z["refine_loss"])) <= 1e-5 * abs(float(z["refine_loss"]))
    assert abs(dl - float(z["det_loss"])) <= 1e-5 * abs(float(z["det_loss"]))
    assert abs(cl["clf_loss"] - float(z["clf_loss"])) <= 1e-5 * abs(float(z["clf_loss"]))
    assert cl["n_pos"] > 0 and cl["n_neg"] == 3 * cl["n_pos"] + 4
