"""GPU parity: the CUDA path (through the Python mirror -> DLPack -> C ABI) against
 (i) the fixtures produced by the UNMODIFIED reference (tests/golden, Tier A) and
 (ii) the NumPy restatement (oracle/restated.py, Tier B) on seeded random inputs.

Bars: integer / index / mask / keep-set outputs bit-exact; floats that involve exp/log within
rtol 1e-5 (atol 1e-6) of the reference fixtures and bit-exact against Tier B (both evaluate
exp/log as the correctly rounded float32)."""
import numpy as np
import pytest
import torch

from conftest import LAYOUTS, golden, golden_anchors, golden_files
from helpers import bit_equal, flat_from_list, to_cuda_list
from oracle import restated as R

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def rb(cuda_device):
    import rodet_b200
    from rodet_b200 import config
    from rodet_b200.utils import common_tools, net_tools
    import rodet_b200.utils.tf_extended as tfe

    class NS:
        pass
    ns = NS()
    ns.pkg, ns.config, ns.nt, ns.ct, ns.tfe, ns.dev = rodet_b200, config, net_tools, common_tools, tfe, cuda_device
    return ns


def _cuda(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ------------------------------------------------------------------------------- anchors
@pytest.mark.parametrize("layout", ["418", "512", "tiny"])
def test_anchor_table_device(rb, layout):
    img, feats = LAYOUTS[layout]
    ref = golden_anchors(layout)
    want = R.AnchorTable(ref)
    # (a) from the reference's anchors_all_layer() output
    t1 = rb.pkg.AnchorTable.from_anchors(ref, rb.dev)
    assert np.array_equal(t1.corner.cpu().numpy(), want.corner)
    assert np.array_equal(t1.center.cpu().numpy(), want.center)
    # (b) generated on the device from init_anchor() parameters (float64 math in the kernel)
    rb.config.img_size = img
    sizes = rb.nt.init_anchor(6)
    assert np.array_equal(np.concatenate(list(sizes.values())), golden("anchors.npz")["%s_sizes_px" % layout])
    t2 = rb.pkg.AnchorTable.generate(img, feats, sizes, rb.dev)
    assert np.array_equal(t2.corner.cpu().numpy(), want.corner)
    assert np.array_equal(t2.center.cpu().numpy(), want.center)
    yxhw = t2.yxhw.cpu().numpy()
    off = 0
    for (y, x, h, w), (fh, fw, a) in zip(ref, want.shapes):
        blk = yxhw[off:off + fh * fw * a].reshape(fh, fw, a, 4)
        assert np.array_equal(blk[..., 0], np.broadcast_to(y, (fh, fw, a)))
        assert np.array_equal(blk[..., 1], np.broadcast_to(x, (fh, fw, a)))
        assert np.array_equal(blk[..., 2], np.broadcast_to(h, (fh, fw, a)))
        assert np.array_equal(blk[..., 3], np.broadcast_to(w, (fh, fw, a)))
        off += fh * fw * a
    # host mirror of anchors_all_layer
    feats_d = {"layer_%d" % (i + 1): f for i, f in enumerate(feats)}
    ours = rb.nt.anchors_all_layer(img, feats_d, sizes)
    for o, r in zip(ours, ref):
        for a, b in zip(o, r):
            assert a.dtype == np.float32 and np.array_equal(a, b)


def test_box_format(rb):
    z = golden("box_format.npz")
    c = rb.ct.cornerBboxes_2_centerBboxes(_cuda(z["corner"], rb.dev))
    assert c.shape == z["center"].shape and np.array_equal(c.cpu().numpy(), z["center"])
    b = rb.ct.centerBboxes_2_cornerBboxes(_cuda(z["center"], rb.dev))
    assert np.array_equal(b.cpu().numpy(), z["corner_back"])


# ------------------------------------------------------------------------------- ARM / ODM / decode vs fixtures
@pytest.mark.parametrize("fname", golden_files("targets_"))
def test_targets_vs_reference_fixture(rb, fname):
    z = golden(fname)
    layout = str(z["layout"])
    anchors = golden_anchors(layout)
    table = R.AnchorTable(anchors)
    shapes = table.shapes
    center = _cuda(z["center"], rb.dev)
    labels = _cuda(z["labels"], rb.dev)
    JB = rb.config.refine_method.JACCARD_BIGGER
    gt, cb, lab, pos, idx = rb.nt.refine_groundtruth(anchors, center, labels, JB, return_match_index=True)
    # reference structure: lists over layers, per image [fh,fw,A,4] / [fh,fw,A,1]
    assert [tuple(t.shape) for t in gt] == [(fh, fw, a, 4) for fh, fw, a in shapes]
    assert [tuple(t.shape) for t in pos] == [(fh, fw, a, 1) for fh, fw, a in shapes]
    assert lab[0].dtype == torch.int32 and pos[0].dtype == torch.int32
    fl = lambda ts, tail: flat_from_list([t.unsqueeze(0) for t in ts], tail)[0]
    g_gt, g_cb, g_lab, g_pos, g_idx = fl(gt, 1), fl(cb, 1), fl(lab, 1)[..., 0], fl(pos, 1)[..., 0], fl(idx, 0)
    assert np.array_equal(g_pos, z["jb_pos"])
    assert np.array_equal(g_lab, z["jb_labels"])
    assert np.array_equal(g_idx, z["jb_idx"])
    o_gt, o_cb, _, _, _ = R.arm_match_encode(table, z["center"], z["labels"])
    assert bit_equal(g_gt, o_gt) and bit_equal(g_cb, o_cb)            # vs Tier B: bit-exact
    if "jb_gt" in z:
        assert bit_equal(g_cb, z["jb_cb"])
        np.testing.assert_allclose(g_gt, z["jb_gt"], rtol=RTOL, atol=ATOL)
    else:
        np.testing.assert_allclose(g_gt[g_pos > 0], z["jb_gt_sparse"], rtol=RTOL, atol=ATOL)
        assert bit_equal(g_cb[g_pos > 0], z["jb_cb_sparse"])
    if "nn_gt" in z:
        NN = rb.config.refine_method.NEAREST_NEIGHBOR
        ngt, ncb, nlab, npos = rb.nt.refine_groundtruth(anchors, center, labels, NN)
        assert np.array_equal(fl(npos, 1)[..., 0], z["nn_pos"])
        assert np.array_equal(fl(nlab, 1)[..., 0], z["nn_labels"])
        assert bit_equal(fl(ncb, 1), z["nn_cb"])
        np.testing.assert_allclose(fl(ngt, 1), z["nn_gt"], rtol=RTOL, atol=ATOL)
    if "enc_gt0" in z:
        enc = [rb.nt.encode_locations_one_layer(a_l, center[0]) for a_l in anchors]
        np.testing.assert_allclose(fl(enc, 1), z["enc_gt0"], rtol=RTOL, atol=ATOL)
    if "odm_refine_out" not in z:
        return
    # ODM fed the REFERENCE's ARM outputs
    B = z["odm_mask"].shape[0]
    n = table.n
    ro = z["odm_refine_out"]
    og = np.broadcast_to(z["jb_gt"], (B, n, 4)).copy()
    cbb = np.broadcast_to(z["jb_cb"], (B, n, 4)).copy()
    lb = np.broadcast_to(z["jb_labels"], (B, n)).copy()
    pm = np.broadcast_to(z["jb_pos"], (B, n)).copy()
    out = rb.nt.det_groundtruth(to_cuda_list(ro, shapes, (4,), rb.dev), to_cuda_list(og, shapes, (4,), rb.dev),
                                to_cuda_list(cbb, shapes, (4,), rb.dev), to_cuda_list(lb, shapes, (1,), rb.dev),
                                to_cuda_list(pm, shapes, (1,), rb.dev), anchors)
    det_gt, mask, dlab, iou = out
    assert [tuple(t.shape) for t in mask] == [(B, fh, fw, a, 1) for fh, fw, a in shapes]
    assert [tuple(t.shape) for t in iou] == [(B, fh, fw, a) for fh, fw, a in shapes]
    assert np.array_equal(flat_from_list(mask, 1)[..., 0], z["odm_mask"])
    assert np.array_equal(flat_from_list(dlab, 1)[..., 0], z["odm_labels"])
    assert bit_equal(flat_from_list(det_gt, 1), z["odm_det_gt"])
    np.testing.assert_allclose(flat_from_list(iou, 0), z["odm_iou"], rtol=1e-4, atol=ATOL)
    o = R.odm_target(table, ro, og, cbb, lb, pm)
    assert bit_equal(flat_from_list(iou, 0), o[3])                     # vs Tier B: bit-exact
    # decode call site
    do = z["dec_det_out"]
    locs = []
    ro_l, do_l = to_cuda_list(ro, shapes, (4,), rb.dev), to_cuda_list(do, shapes, (4,), rb.dev)
    for a_l, r_l, d_l in zip(anchors, ro_l, do_l):
        c = rb.nt.decode_locations_one_layer(a_l, r_l + d_l)
        assert c.shape == r_l.shape
        locs.append(rb.ct.centerBboxes_2_cornerBboxes(c))
    got = flat_from_list(locs, 1)
    np.testing.assert_allclose(got, z["dec_corner"], rtol=RTOL, atol=ATOL)
    assert bit_equal(got, R.decode_corner(table, ro, do))


# ------------------------------------------------------------------------------- post-process vs fixtures
def _detect_inputs(z, table):
    if "probs" in z:
        return z["probs"], z["refine_out"], z["det_out"]
    from test_oracle_golden import _regen_detect_inputs
    return _regen_detect_inputs(z, table)


@pytest.mark.parametrize("fname", golden_files("detect_"))
def test_detected_bboxes_vs_reference_fixture(rb, fname):
    z = golden(fname)
    layout = str(z["layout"])
    anchors = golden_anchors(layout)
    table = R.AnchorTable(anchors)
    shapes = table.shapes
    probs, ro, do = _detect_inputs(z, table)
    sthr = None if bool(z["select_none"]) else round(float(z["select_threshold"]), 6)
    nthr, topk, keep = round(float(z["nms_threshold"]), 6), int(z["top_k"]), int(z["keep_top_k"])
    clip = z["clip"] if "clip" in z else None
    preds = to_cuda_list(probs, shapes, (11,), rb.dev)
    ro_l, do_l = to_cuda_list(ro, shapes, (4,), rb.dev), to_cuda_list(do, shapes, (4,), rb.dev)
    # (1) drop-in call sequence with the reference's own decoded boxes (bit-exact stage parity)
    if "boxes_corner" in z:
        locs = to_cuda_list(z["boxes_corner"], shapes, (4,), rb.dev)
        rs, rbx = rb.nt.detected_bboxes(preds, locs, select_threshold=sthr, nms_threshold=nthr,
                                        clipping_bbox=clip, top_k=topk, keep_top_k=keep)
        assert sorted(rs.keys()) == list(range(1, 11))
        for c in range(1, 11):
            assert rs[c].shape == (probs.shape[0], keep) and rbx[c].shape == (probs.shape[0], keep, 4)
            assert bit_equal(rs[c].cpu().numpy(), z["scores_c%d" % c]), c
            assert bit_equal(rbx[c].cpu().numpy(), z["bboxes_c%d" % c]), c
        # stage outputs: select, sort
        d_s, d_b = rb.nt.bboxes_select_all_layers(preds, locs, select_threshold=sthr, num_classes=11)
        s_s, s_b = rb.tfe.bboxes_sort(d_s, d_b, top_k=topk)
        n_s, n_b = rb.tfe.bboxes_nms_batch(s_s, s_b, nms_threshold=nthr, keep_top_k=keep)
        if clip is not None:
            n_b = rb.tfe.bboxes_clip(clip, n_b)
        for c in range(1, 11):
            assert bit_equal(d_s[c].cpu().numpy(), z["sel_scores_c%d" % c])
            assert bit_equal(d_b[c].cpu().numpy(), z["sel_bboxes_c%d" % c])
            assert bit_equal(s_s[c].cpu().numpy(), z["sort_scores_c%d" % c])
            assert bit_equal(s_b[c].cpu().numpy(), z["sort_bboxes_c%d" % c])
            assert bit_equal(n_s[c].cpu().numpy(), z["scores_c%d" % c])
            assert bit_equal(n_b[c].cpu().numpy(), z["bboxes_c%d" % c])
    # (2) fully fused decode + post-process
    rs, rbx, counts = rb.nt.decode_detected_bboxes(anchors, ro_l, do_l, preds, select_threshold=sthr,
                                                   nms_threshold=nthr, clipping_bbox=clip, top_k=topk,
                                                   keep_top_k=keep, return_counts=True)
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), sthr, nthr, clip, topk, keep)
    for c in range(1, 11):
        got_s, got_b = rs[c].cpu().numpy(), rbx[c].cpu().numpy()
        assert bit_equal(got_s, z["scores_c%d" % c]), c                  # keep set / order vs reference
        np.testing.assert_allclose(got_b, z["bboxes_c%d" % c], rtol=RTOL, atol=ATOL)
        assert bit_equal(got_s, o_s[c]) and bit_equal(got_b, o_b[c])     # vs Tier B: bit-exact
        assert np.array_equal(counts[c].cpu().numpy(), (o_s[c] != 0).sum(axis=1))


def test_tfe_ops_vs_reference_fixture(rb):
    z = golden("tfe_ops.npz")
    b, r, s = _cuda(z["boxes"], rb.dev), _cuda(z["ref"], rb.dev), _cuda(z["scores"], rb.dev)
    assert bit_equal(rb.tfe.bboxes_jaccard(r, b).cpu().numpy(), z["jaccard"])
    assert bit_equal(rb.tfe.bboxes_intersection(r, b).cpu().numpy(), z["intersection"])
    assert bit_equal(rb.tfe.bboxes_resize(r, b).cpu().numpy(), z["resize"])
    assert bit_equal(rb.tfe.bboxes_clip(r, b).cpu().numpy(), z["clip"])
    ns, nb = rb.tfe.bboxes_nms(s, b, nms_threshold=0.3, keep_top_k=25)
    assert bit_equal(ns.cpu().numpy(), z["nms_scores"]) and bit_equal(nb.cpu().numpy(), z["nms_bboxes"])
    ss, sb = rb.tfe.bboxes_sort(s[None], b[None], top_k=20)
    assert bit_equal(ss.cpu().numpy(), z["sort_scores"]) and bit_equal(sb.cpu().numpy(), z["sort_bboxes"])
    assert bit_equal(rb.tfe.pad_axis(b[:5], 0, 9, axis=0).cpu().numpy(), z["pad"])
    # net_tools.jaccard (plain divide) against the oracle
    a = R.AnchorTable(golden_anchors("tiny")).corner
    g = z["boxes"][0]
    assert bit_equal(rb.nt.jaccard(_cuda(a, rb.dev), _cuda(g, rb.dev)).cpu().numpy(), R.jaccard(a, g))
    assert bit_equal(rb.nt.jaccard(_cuda(z["boxes"], rb.dev), _cuda(z["boxes"][::-1].copy(), rb.dev)).cpu().numpy(),
                     R.jaccard(z["boxes"], z["boxes"][::-1]))


def test_eval_matching_vs_reference_fixture(rb):
    """f-2: TP / FP matching (tfe.bboxes_matching_batch) against the reference fixture, dict and tensor forms."""
    z = golden("eval_matching.npz")
    cl = [int(c) for c in z["classes"]]
    sc = {c: _cuda(z["scores_c%d" % c], rb.dev) for c in cl}
    bx = {c: _cuda(z["bboxes_c%d" % c], rb.dev) for c in cl}
    gl, gb, gd = _cuda(z["glabels"], rb.dev), _cuda(z["gbboxes"], rb.dev), _cuda(z["gdifficults"], rb.dev)
    n, tp, fp, s2 = rb.tfe.bboxes_matching_batch(cl, sc, bx, gl, gb, gd, matching_threshold=float(z["thr"]))
    assert s2 is sc
    for c in cl:
        assert n[c].dtype == torch.int64 and tp[c].dtype == torch.bool
        assert np.array_equal(n[c].cpu().numpy(), z["n_c%d" % c])
        assert np.array_equal(tp[c].cpu().numpy(), z["tp_c%d" % c])
        assert np.array_equal(fp[c].cpu().numpy(), z["fp_c%d" % c])
    c = cl[0]
    n1, tp1, fp1 = rb.tfe.bboxes_matching(c, sc[c][1], bx[c][1], gl[1], gb[1], gd[1], matching_threshold=float(z["thr"]))
    assert int(n1) == int(z["n_c%d" % c][1]) and np.array_equal(tp1.cpu().numpy(), z["tp_c%d" % c][1])
    assert np.array_equal(fp1.cpu().numpy(), z["fp_c%d" % c][1])
    # int32 ground-truth labels and a larger randomised case against the NumPy restatement
    rng = np.random.default_rng(5)
    B, N, G = 5, 200, 70
    g_b = np.sort(rng.uniform(0, 1, size=(B, G, 2, 2)).astype(np.float32), axis=2).reshape(B, G, 4)
    g_l = rng.integers(0, 4, size=(B, G)).astype(np.int32)
    g_d = (rng.uniform(size=(B, G)) < 0.15).astype(np.int32)
    d_b = (g_b[:, rng.integers(0, G, size=N)] + rng.normal(0, 0.01, size=(B, N, 4))).astype(np.float32)
    d_s = np.sort(rng.uniform(0, 1, size=(B, N)).astype(np.float32), axis=1)[:, ::-1].copy()
    n2, tp2, fp2, _ = rb.tfe.bboxes_matching_batch(2, _cuda(d_s, rb.dev), _cuda(d_b, rb.dev), _cuda(g_l, rb.dev),
                                                   _cuda(g_b, rb.dev), _cuda(g_d, rb.dev))
    o = R.bboxes_matching_batch(2, d_s, d_b, g_l, g_b, g_d)
    assert np.array_equal(n2.cpu().numpy(), o[0]) and np.array_equal(tp2.cpu().numpy(), o[1])
    assert np.array_equal(fp2.cpu().numpy(), o[2]) and o[1].sum() > 20


def test_cascade_decode_extra(cuda_device):
    """Opt-in RefineDet cascade decode (no reference counterpart; BASELINE north star): bit-exact against the restated
    composition of two reference decodes, different from the reference's sum-then-decode, and usable as the
    `localisations` of detected_bboxes."""
    from rodet_b200 import synth
    from rodet_b200.utils import net_tools
    anchors = golden_anchors("418")
    table = R.AnchorTable(anchors)
    B = 2
    ro = np.stack([synth.head_offsets(b, table.n, 0) for b in range(B)])
    do = np.stack([synth.head_offsets(b, table.n, 1) for b in range(B)])
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), cuda_device), to_cuda_list(do, table.shapes, (4,), cuda_device)
    out = net_tools.decode_locations_cascade(anchors, ro_l, do_l)
    got = flat_from_list(out, 1)
    assert bit_equal(got, R.decode_cascade_corner(table, ro, do))
    assert not np.allclose(got, R.decode_corner(table, ro, do))            # not the reference's single decode
    probs = np.stack([synth.class_probs(b, table.n) for b in range(B)])
    rs, rb = net_tools.detected_bboxes(to_cuda_list(probs, table.shapes, (11,), cuda_device), out, select_threshold=0.3,
                                       nms_threshold=0.45, top_k=400, keep_top_k=200)
    o_s, o_b = R.detected_bboxes(probs, got, 0.3, 0.45, None, 400, 200)
    for c in range(1, 11):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]) and bit_equal(rb[c].cpu().numpy(), o_b[c])


@pytest.mark.parametrize("layout,first,B", [("418", 60, 3), ("512", 64, 2)])
def test_forced_match_extra(cuda_device, layout, first, B):
    """Opt-in forced match (no reference counterpart; BASELINE north star): every GT box claims its best anchor.  Against
    the restated definition, bit-exact; the default call is unchanged; a GT no anchor reaches the threshold for gains a
    positive."""
    from rodet_b200 import config, synth
    from rodet_b200.utils import net_tools
    anchors = golden_anchors(layout)
    table = R.AnchorTable(anchors)
    corner, labels, counts = synth.gt_batch(first, B, max_gt=30)
    corner[0, 0] = np.array([0.40, 0.40, 0.404, 0.401], np.float32)      # a sliver no anchor matches by threshold
    center = R.corner_to_center(corner).astype(np.float32)
    for b in range(B):
        center[b, counts[b]:] = 0
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    JB = config.refine_method.JACCARD_BIGGER
    base = net_tools.refine_groundtruth(anchors, d(center), d(labels), JB, gt_counts=d(counts), return_match_index=True)
    forced = net_tools.refine_groundtruth(anchors, d(center), d(labels), JB, gt_counts=d(counts), return_match_index=True,
                                          forced_match=True)
    fl = lambda ts, k: [flat_from_list(x, t) for x, t in zip(ts, k)]
    gb, gf = fl(base, (1, 1, 1, 1, 0)), fl(forced, (1, 1, 1, 1, 0))
    gained = 0
    for b in range(B):
        k = int(counts[b])
        o = R.arm_match_encode(table, center[b, :k], labels[b, :k])
        assert np.array_equal(gb[3][b, :, 0], o[3]) and np.array_equal(gb[4][b], o[4])          # default path unchanged
        f = R.forced_match(table, center[b, :k], labels[b, :k], o)
        assert np.array_equal(gf[3][b, :, 0], f[3]) and np.array_equal(gf[4][b], f[4]) and np.array_equal(gf[2][b, :, 0], f[2])
        assert bit_equal(gf[1][b], f[1]) and bit_equal(gf[0][b], f[0])
        gained += int(f[3].sum() - o[3].sum())
        claimed = set(int(v) for v in f[4][f[3] > 0])
        best = R.jaccard(table.corner[None], R.center_to_corner(center[b, :k])[:, None]).argmax(1)
        for g in range(k):            # a GT without any positive lost its best anchor to a rival GT that claims the same anchor
            assert g in claimed or any(o2 != g and best[o2] == best[g] for o2 in range(k))
    assert gained > 0
    with pytest.raises(ValueError):
        net_tools.refine_groundtruth(anchors, d(center), d(labels), config.refine_method.NEAREST_NEIGHBOR, gt_counts=d(counts),
                                     forced_match=True)
