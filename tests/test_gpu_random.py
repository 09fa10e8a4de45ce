"""GPU parity on the BASELINE-shaped synthetic workload against Tier B,
at sizes the NumPy oracle finishes in seconds, plus size-independent properties at the full
BASELINE batch sizes, plus the 1k-image bit-exactness run the north star asks for."""
import os

import numpy as np
import pytest
import torch

from conftest import LAYOUTS, golden_anchors
from helpers import bit_equal, flat_from_list, to_cuda_list
from oracle import restated as R

pytestmark = pytest.mark.gpu
N_PARITY_IMAGES = int(os.environ.get("RODET_PARITY_IMAGES", "1000"))


@pytest.fixture(scope="module")
def env(cuda_device):
    import rodet_b200
    from rodet_b200 import config, synth
    from rodet_b200.utils import common_tools, net_tools

    class NS:
        pass
    ns = NS()
    ns.pkg, ns.config, ns.synth, ns.nt, ns.ct, ns.dev = rodet_b200, config, synth, net_tools, common_tools, cuda_device
    ns.anchors = {k: golden_anchors(k) for k in ("418", "512")}
    ns.otable = {k: R.AnchorTable(v) for k, v in ns.anchors.items()}
    return ns


def _gt_center(env, first, B):
    corner, labels, counts = env.synth.gt_batch(first, B)
    center = R.corner_to_center(corner)
    for b in range(B):
        center[b, counts[b]:] = 0
    return corner, center.astype(np.float32), labels, counts


def _arm_gpu(env, layout, center, labels, counts):
    JB = env.config.refine_method.JACCARD_BIGGER
    out = env.nt.refine_groundtruth(env.anchors[layout], torch.from_numpy(center).to(env.dev),
                                    torch.from_numpy(labels).to(env.dev), JB,
                                    gt_counts=torch.from_numpy(counts).to(env.dev), return_match_index=True)
    return out


def _check_arm(env, layout, first, B):
    table = env.otable[layout]
    corner, center, labels, counts = _gt_center(env, first, B)
    # the host converts corner -> centre with the library too (train.py:109)
    cen_gpu = env.ct.cornerBboxes_2_centerBboxes(torch.from_numpy(corner).to(env.dev)).cpu().numpy()
    for b in range(B):
        assert bit_equal(cen_gpu[b, :counts[b]], center[b, :counts[b]])
    gt, cb, lab, pos, idx = _arm_gpu(env, layout, center, labels, counts)
    g = [flat_from_list(x, t) for x, t in ((gt, 1), (cb, 1), (lab, 1), (pos, 1), (idx, 0))]
    npos = 0
    for b in range(B):
        o = R.arm_match_encode(table, center[b, :counts[b]], labels[b, :counts[b]])
        assert np.array_equal(g[3][b, :, 0], o[3]), "pos mask, image %d" % (first + b)
        assert np.array_equal(g[4][b], o[4]), "match index, image %d" % (first + b)
        assert np.array_equal(g[2][b, :, 0], o[2]), "labels, image %d" % (first + b)
        assert bit_equal(g[1][b], o[1]) and bit_equal(g[0][b], o[0])
        npos += int(o[3].sum())
    return (gt, cb, lab, pos), g, npos


@pytest.mark.parametrize("layout", ["512", "418"])
def test_arm_odm_random_batch(env, layout):
    B = 6
    table = env.otable[layout]
    (gt, cb, lab, pos), g, npos = _check_arm(env, layout, 100, B)
    assert npos > 0
    # ODM on the GPU's own ARM outputs (zero-copy per-layer views) + synthetic ARM head output
    ro = np.stack([env.synth.head_offsets(100 + b, table.n) for b in range(B)])
    # move half of the positives close to their target so both mask values occur
    close = (np.arange(table.n)[None, :, None] % 2 == 0) & (g[3] > 0)
    ro = np.where(close, g[0] + 0.05 * ro, ro).astype(np.float32)
    ro_l = to_cuda_list(ro, table.shapes, (4,), env.dev)
    det_gt, mask, dlab, iou = env.nt.det_groundtruth(ro_l, gt, cb, lab, pos, env.anchors[layout])
    o = R.odm_target(table, ro, g[0], g[1], g[2][..., 0], g[3][..., 0])
    assert np.array_equal(flat_from_list(mask, 1)[..., 0], o[1])
    assert np.array_equal(flat_from_list(dlab, 1)[..., 0], o[2])
    assert bit_equal(flat_from_list(iou, 0), o[3])
    assert bit_equal(flat_from_list(det_gt, 1), o[0])
    assert 0 < int(o[1].sum()) < int(g[3].sum())


def _detect_inputs(env, layout, first, B, stress):
    table = env.otable[layout]
    mk = env.synth.stress_probs if stress else env.synth.class_probs
    probs = np.stack([mk(first + b, table.n) for b in range(B)])
    s = (0.05, 0.05) if stress else (0.1, 0.2)
    ro = np.stack([env.synth.head_offsets(first + b, table.n, 0, *s) for b in range(B)])
    do = np.stack([env.synth.head_offsets(first + b, table.n, 1, *s) for b in range(B)])
    return probs, ro, do


def _check_detect(env, layout, first, B, stress, sthr=0.3, nthr=0.45, topk=400, keep=200):
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, first, B, stress)
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    rs, rb, counts = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=sthr,
                                                   nms_threshold=nthr, top_k=topk, keep_top_k=keep, return_counts=True)
    boxes = R.decode_corner(table, ro, do)
    o_s, o_b = R.detected_bboxes(probs, boxes, sthr, nthr, None, topk, keep)
    ndet = 0
    for c in range(1, 11):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]), "scores class %d images %d.." % (c, first)
        assert bit_equal(rb[c].cpu().numpy(), o_b[c]), "boxes class %d images %d.." % (c, first)
        ndet += int((o_s[c] != 0).sum())
    # drop-in sequence: decode per layer -> c2c -> detected_bboxes must give the same
    locs = [env.ct.centerBboxes_2_cornerBboxes(env.nt.decode_locations_one_layer(a, r + d))
            for a, r, d in zip(env.anchors[layout], ro_l, do_l)]
    assert bit_equal(flat_from_list(locs, 1), boxes)
    rs2, rb2 = env.nt.detected_bboxes(preds, locs, select_threshold=sthr, nms_threshold=nthr, top_k=topk, keep_top_k=keep)
    for c in range(1, 11):
        assert torch.equal(rs2[c], rs[c]) and bit_equal(rb2[c].cpu().numpy(), rb[c].cpu().numpy())
    return ndet


@pytest.mark.parametrize("layout,stress", [("512", False), ("512", True), ("418", False)])
def test_detect_random_batch(env, layout, stress):
    assert _check_detect(env, layout, 200, 3, stress) > 0


def test_detect_script_defaults(env):
    """evaluate.py:58-65 (0.3 / 0.4 / 400 / 200), predict.py:136-137 (0.1 / 0.4) and the
    signature defaults of detected_bboxes (None / 0.5 / 800 / 200)."""
    _check_detect(env, "418", 300, 1, False, 0.3, 0.4, 400, 200)
    _check_detect(env, "418", 301, 1, False, 0.1, 0.4, 400, 200)
    _check_detect(env, "418", 302, 1, False, None, 0.5, 800, 200)


def test_thousand_images_bit_exact(env):
    """North star: match indices, positive masks and NMS keep sets bit-exact on 1k synthetic
    images (512x512 layout).  RODET_PARITY_IMAGES shortens the run."""
    n = N_PARITY_IMAGES
    step = 8
    npos = ndet = 0
    for first in range(0, n, step):
        B = min(step, n - first)
        _, _, p = _check_arm(env, "512", 10_000 + first, B)
        npos += p
    for first in range(0, n, 4):
        ndet += _check_detect(env, "512", 20_000 + first, min(4, n - first), stress=(first % 40 == 0))
    assert npos > n and ndet > n


# ------------------------------------------------------------------------------- full-size properties
def test_fullsize_properties_match_encode(env):
    """BASELINE config 2 (batch 32): properties that need no oracle."""
    B, layout = 32, "512"
    table = env.otable[layout]
    corner, center, labels, counts = _gt_center(env, 5000, B)
    gt, cb, lab, pos, idx = _arm_gpu(env, layout, center, labels, counts)
    P = flat_from_list(pos, 1)[..., 0]
    I = flat_from_list(idx, 0)
    L = flat_from_list(lab, 1)[..., 0]
    C = flat_from_list(cb, 1)
    G = flat_from_list(gt, 1)
    assert set(np.unique(P)) <= {0, 1}
    assert (I >= 0).all() and (I < counts[:, None]).all()
    for b in range(B):
        m = P[b] > 0
        assert bit_equal(C[b][m], center[b][I[b][m]])             # matched box is the argmax box
        assert np.array_equal(L[b][m], labels[b][I[b][m]].astype(np.int32))
        assert not C[b][~m].any() and not G[b][~m].any() and not L[b][~m].any()
    # decode(encode(gt)) round trip recovers the matched boxes on positives
    dec = [env.nt.decode_locations_one_layer(a, g) for a, g in zip(env.anchors[layout], gt)]
    D = flat_from_list(dec, 1)
    m = P > 0
    np.testing.assert_allclose(D[m], C[m], rtol=2e-5, atol=2e-6)
    # idempotence: a second launch gives identical bits
    gt2 = _arm_gpu(env, layout, center, labels, counts)[0]
    assert all(torch.equal(a, b) for a, b in zip(gt, gt2))


def test_fullsize_properties_detect(env):
    """BASELINE configs 3 and 5 (batch 64, normal and stress scores)."""
    layout, B, keep, nthr = "512", 64, 200, 0.45
    table = env.otable[layout]
    for stress in (False, True):
        probs, ro, do = _detect_inputs(env, layout, 7000, B, stress)
        preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
        ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
        rs, rb, counts = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.3,
                                                       nms_threshold=nthr, top_k=400, keep_top_k=keep, return_counts=True)
        for c in (1, 5, 10):
            s, b = rs[c].cpu().numpy(), rb[c].cpu().numpy()
            assert (np.diff(s, axis=1) <= 0).all()                      # descending, zero padded at the end
            cnt = counts[c].cpu().numpy()
            assert np.array_equal(cnt, (s != 0).sum(1))
            assert ((s >= np.float32(0.3)) | (s == 0)).all()
            for img in (0, B - 1):
                k = cnt[img]
                assert not b[img, k:].any()
                iou = R.nms_iou_matrix(b[img, :k])
                np.fill_diagonal(iou, 0)
                assert (iou <= np.float32(nthr)).all()                  # survivors do not suppress each other
        if stress:
            assert int(counts[1:].sum()) >= int(0.9 * 10 * B * keep)


def test_detect_heavy_ties_overflow_fallback(env):
    """Scores quantised to 1/8: thousands of exactly equal scores per class, so the streaming
    candidate list overflows and the exact general kernels must take over (lowest-index ties)."""
    layout, B = "512", 2
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, 400, B, stress=True)
    probs = (np.round(probs * 8) / 8).astype(np.float32)
    probs[1, :, 3] = np.float32(0.5)                      # one class entirely tied
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    rs, rb, counts = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.3,
                                                   nms_threshold=0.45, top_k=400, keep_top_k=200, return_counts=True)
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), 0.3, 0.45, None, 400, 200)
    for c in range(1, 11):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]), c
        assert bit_equal(rb[c].cpu().numpy(), o_b[c]), c
        assert np.array_equal(counts[c].cpu().numpy(), (o_s[c] != 0).sum(1))


def test_detect_fewer_candidates_than_topk(env):
    """High threshold: fewer than top_k candidates per class, some classes empty."""
    layout, B = "512", 2
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, 410, B, stress=False)
    probs[0, :, 4] = np.float32(0.01)                     # class 4 of image 0: no candidate at all
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    rs, rb = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.9,
                                           nms_threshold=0.45, top_k=400, keep_top_k=200)
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), 0.9, 0.45, None, 400, 200)
    n = 0
    for c in range(1, 11):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]) and bit_equal(rb[c].cpu().numpy(), o_b[c]), c
        n += int((o_s[c] != 0).sum())
    assert n > 0 and not rs[4][0].any()


def test_detect_generic_class_count(env, monkeypatch):
    """Prediction depth != 11 takes the plain-load two-pass kernels (no TMA tiles)."""
    layout, B, C = "418", 2, 6
    table = env.otable[layout]
    rng = np.random.default_rng(77)
    z = (rng.standard_normal(size=(B, table.n, C)) * 3.0).astype(np.float32)
    z[..., 0] += np.float32(2.0)
    z -= z.max(-1, keepdims=True)
    probs = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    ro = np.stack([env.synth.head_offsets(600 + b, table.n) for b in range(B)])
    do = np.stack([env.synth.head_offsets(600 + b, table.n, 1) for b in range(B)])
    preds = to_cuda_list(probs, table.shapes, (C,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    monkeypatch.setattr(env.config, "total_obj_n", C)
    rs, rb = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.3,
                                           nms_threshold=0.45, top_k=400, keep_top_k=200)
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), 0.3, 0.45, None, 400, 200, num_classes=C)
    assert sorted(rs.keys()) == list(range(1, C))
    for c in range(1, C):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]) and bit_equal(rb[c].cpu().numpy(), o_b[c]), c


def test_detect_large_topk_and_small_threshold(env):
    """top_k = 800 (signature default, sort capacity 2048) with a low threshold: many candidates."""
    assert _check_detect(env, "418", 700, 2, False, 0.05, 0.5, 800, 300) > 0
    assert _check_detect(env, "512", 702, 1, True, 0.3, 0.45, 1000, 250) > 0


def test_detect_batch_extremes(env):
    """One image (64 scanning CTAs would be too many: clamped) and a large batch (256 images)."""
    assert _check_detect(env, "512", 710, 1, False) > 0
    layout, B = "418", 256
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, 720, 8, False)
    probs, ro, do = (np.tile(a, (B // 8,) + (1,) * (a.ndim - 1)) for a in (probs, ro, do))
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    rs, rb = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.3,
                                           nms_threshold=0.45, top_k=400, keep_top_k=200)
    o_s, o_b = R.detected_bboxes(probs[:8], R.decode_corner(table, ro[:8], do[:8]), 0.3, 0.45, None, 400, 200)
    for c in range(1, 11):
        got_s, got_b = rs[c].cpu().numpy(), rb[c].cpu().numpy()
        for rep in (0, 13, B // 8 - 1):
            assert bit_equal(got_s[rep * 8:(rep + 1) * 8], o_s[c]) and bit_equal(got_b[rep * 8:(rep + 1) * 8], o_b[c])
    # empty batch
    e = lambda ts: [t[:0] for t in ts]
    rs0, rb0 = env.nt.decode_detected_bboxes(env.anchors[layout], e(ro_l), e(do_l), e(preds), select_threshold=0.3)
    assert rs0[1].shape == (0, 200) and rb0[1].shape == (0, 200, 4)


def test_arm_gt_count_extremes(env):
    """A single GT box, and more GT boxes than one 256-thread staging pass (gmax = 300)."""
    layout = "418"
    table = env.otable[layout]
    JB = env.config.refine_method.JACCARD_BIGGER
    rng = np.random.default_rng(3)
    for G in (1, 300):
        c = rng.uniform(0.1, 0.9, size=(2, G, 2))
        hw = np.exp(rng.uniform(np.log(0.03), np.log(0.5), size=(2, G, 2)))
        corner = np.clip(np.concatenate([c - hw / 2, c + hw / 2], -1), 0, 1).astype(np.float32)
        center = R.corner_to_center(corner).astype(np.float32)
        labels = rng.integers(1, 11, size=(2, G)).astype(np.int64)
        counts = np.array([G, max(1, G // 2)], np.int32)
        out = env.nt.refine_groundtruth(env.anchors[layout], torch.from_numpy(center).to(env.dev),
                                        torch.from_numpy(labels).to(env.dev), JB,
                                        gt_counts=torch.from_numpy(counts).to(env.dev), return_match_index=True)
        g = [flat_from_list(x, t) for x, t in zip(out, (1, 1, 1, 1, 0))]
        for b in range(2):
            o = R.arm_match_encode(table, center[b, :counts[b]], labels[b, :counts[b]])
            assert np.array_equal(g[3][b, :, 0], o[3]) and np.array_equal(g[4][b], o[4])
            assert bit_equal(g[0][b], o[0]) and bit_equal(g[1][b], o[1]) and np.array_equal(g[2][b, :, 0], o[2])
    # per-image (reference) form == batched form
    one = env.nt.refine_groundtruth(env.anchors[layout], torch.from_numpy(center[0]).to(env.dev),
                                    torch.from_numpy(labels[0]).to(env.dev), JB)
    assert tuple(one[0][0].shape) == table.shapes[0] + (4,)
    assert bit_equal(flat_from_list([t.unsqueeze(0) for t in one[0]], 1)[0], g[0][0])


@pytest.mark.parametrize("where", ["sampled_only", "unsampled_only"])
def test_detect_unrepresentative_sample(env, where):
    """The scan pass cuts candidates below scores estimated from every 19th anchor (n % 19 == 9).
    Put one class's candidates only ONTO those anchors (the estimate overshoots: fewer than top_k survive the cut) or only
    OFF them (no estimate at all): the result must still be exact (general kernels / uncut lists)."""
    layout, B = "512", 2
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, 800, B, stress=False)
    sampled = (np.arange(table.n) % 19) == 9
    rng = np.random.default_rng(5)
    col = np.full((B, table.n), 0.01, np.float32)
    sel = sampled if where == "sampled_only" else ~sampled
    col[:, sel] = rng.uniform(0.3, 1.0, size=(B, int(sel.sum()))).astype(np.float32)
    probs[:, :, 3] = col
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    rs, rb, counts = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=0.3,
                                                   nms_threshold=0.45, top_k=400, keep_top_k=200, return_counts=True)
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), 0.3, 0.45, None, 400, 200)
    for c in range(1, 11):
        assert bit_equal(rs[c].cpu().numpy(), o_s[c]), c
        assert bit_equal(rb[c].cpu().numpy(), o_b[c]), c
        assert np.array_equal(counts[c].cpu().numpy(), (o_s[c] != 0).sum(1))
    assert (o_s[3] != 0).sum() >= B * 150


# ------------------------------------------------------------------------------- bench.py's exact launch geometry
def _detect_vs_c_oracle(env, layout, probs, ro, do, sthr=0.3, nthr=0.45, topk=400, keep=200, ws=None, calls=1):
    """decode_detected_bboxes on the whole batch in ONE launch vs the C oracle (oracle/c, pinned bit-for-bit to
    oracle/restated.py by tests/test_oracle_c.py), every image, scores and boxes bit-exact.  Returns
    (#detections, fallback flags [C,B])."""
    from oracle import c_port as CP
    table = env.otable[layout]
    B = probs.shape[0]
    preds = to_cuda_list(probs, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    if ws is None:
        ws = env.nt.detect_workspace(env.anchors[layout], B, topk, env.dev)
    for _ in range(calls):
        rs, rb, counts = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, preds, select_threshold=sthr,
                                                       nms_threshold=nthr, top_k=topk, keep_top_k=keep, return_counts=True,
                                                       workspace=ws)
    flags = env.nt.detect_fallback_flags(ws).cpu().numpy()
    o_s, o_b = CP.detected_bboxes(probs, CP.decode_corner(table, ro, do), sthr, nthr, topk, keep)
    ndet = 0
    for c in range(1, 11):
        g_s, g_b = rs[c].cpu().numpy(), rb[c].cpu().numpy()
        for b in range(B):
            assert bit_equal(g_s[b], o_s[c][b]), "scores class %d image %d" % (c, b)
            assert bit_equal(g_b[b], o_b[c][b]), "boxes class %d image %d" % (c, b)
        ndet += int((o_s[c] != 0).sum())
        assert np.array_equal(counts[c].cpu().numpy(), (o_s[c] != 0).sum(1))
    return ndet, flags


@pytest.mark.parametrize("stress", [False, True])
def test_detect_bench_geometry_bit_exact(env, stress):
    """BASELINE configs[2] / configs[4] exactly as bench.py launches them: B = 64 images at 512x512 in one call
    (the scan grid, list-slice size and segment grid all depend on the batch), compared with the oracle image by
    image."""
    probs, ro, do = _detect_inputs(env, "512", 500_000, 64, stress)          # bench.py's first decode_nms batch
    ndet, flags = _detect_vs_c_oracle(env, "512", probs, ro, do)
    assert ndet > 64 * 100
    assert flags[1:].mean() < 0.05, "fallback rate on the i.i.d. bench workload: %.3f" % flags[1:].mean()


@pytest.mark.parametrize("fill", ["ones", "random"])
def test_detect_dirty_workspace_costs_speed_not_results(env, fill):
    """The library keeps the bookkeeping head of a workspace clean between calls instead of clearing it per call
    (detect_workspace() zeroes it once).  A workspace that never was — every byte set, or random bytes — must still give
    the oracle's result on the first call (its sampled histogram is garbage: segments go to the exact kernels), and is
    clean from the second call on."""
    probs, ro, do = _detect_inputs(env, "512", 777_000, 4, False)
    ws = env.nt.detect_workspace(env.anchors["512"], 4, 400, env.dev)
    if fill == "ones":
        ws.fill_(0xFF)
    else:
        ws.copy_(torch.randint(0, 256, ws.shape, dtype=torch.uint8, device=env.dev, generator=torch.Generator(env.dev).manual_seed(5)))
    ws._rod_key = (4, 11, 400)
    _detect_vs_c_oracle(env, "512", probs, ro, do, ws=ws)
    ndet, flags = _detect_vs_c_oracle(env, "512", probs, ro, do, ws=ws, calls=2)
    assert ndet > 0 and flags[1:].sum() == 0


def test_detect_default_workspace_and_graph_capture(env):
    """Calls without `workspace=` share one cached workspace per (stream, geometry); under CUDA-graph capture the
    workspace is allocated inside the graph.  Eager calls, repeated calls and graph replays (with the inputs changed in
    place between replays) all give the same bits as a call with an explicit workspace."""
    from rodet_b200.utils import net_tools as NT
    table = env.otable["512"]
    kw = dict(select_threshold=0.3, nms_threshold=0.45, top_k=400, keep_top_k=200)
    sets = []
    for first in (880_000, 881_000):
        probs, ro, do = _detect_inputs(env, "512", first, 3, False)
        sets.append((to_cuda_list(probs, table.shapes, (11,), env.dev), to_cuda_list(ro, table.shapes, (4,), env.dev),
                     to_cuda_list(do, table.shapes, (4,), env.dev)))
    ref = []
    for p, r, d in sets:
        ws = env.nt.detect_workspace(env.anchors["512"], 3, 400, env.dev)
        ref.append(env.nt.decode_detected_bboxes(env.anchors["512"], r, d, p, workspace=ws, **kw))
    same = lambda a, b: all(torch.equal(a[0][c].view(torch.int32), b[0][c].view(torch.int32)) and
                            torch.equal(a[1][c].view(torch.int32), b[1][c].view(torch.int32)) for c in range(1, 11))
    n_cached = len(NT._WS_CACHE)
    for _ in range(2):
        for (p, r, d), want in zip(sets, ref):
            assert same(env.nt.decode_detected_bboxes(env.anchors["512"], r, d, p, **kw), want)
    assert len(NT._WS_CACHE) <= n_cached + 1                       # one workspace for the four calls
    # graph: static input buffers, contents swapped between replays
    p, r, d = [[t.clone() for t in ts] for ts in sets[0]]
    side = torch.cuda.Stream(env.dev)
    side.wait_stream(torch.cuda.current_stream(env.dev))
    with torch.cuda.stream(side):
        env.nt.decode_detected_bboxes(env.anchors["512"], r, d, p, **kw)      # warm-up on the capture stream
    torch.cuda.current_stream(env.dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = env.nt.decode_detected_bboxes(env.anchors["512"], r, d, p, **kw)
    for rep in range(3):
        src = sets[rep % 2]
        for dst_l, src_l in zip((p, r, d), src):
            for a, b in zip(dst_l, src_l):
                a.copy_(b)
        g.replay()
        torch.cuda.synchronize(env.dev)
        assert same(out, ref[rep % 2]), "graph replay %d" % rep


@pytest.mark.parametrize("mode", ["quadrant", "bumps"])
@pytest.mark.parametrize("B", [64, 3])
def test_detect_clustered_scores_bit_exact(env, mode, B):
    """Spatially clustered candidates (a trained head, unlike i.i.d. scores): a class's candidates packed into one
    quadrant of one layer / around the GT boxes.  Whatever route the segments take (sampled cut, list overflow,
    exact fallback kernels) the result is the oracle's."""
    table = env.otable["512"]
    first = 610_000
    probs = np.stack([env.synth.clustered_probs(first + b, table.shapes, mode) for b in range(B)])
    ro = np.stack([env.synth.head_offsets(first + b, table.n, 0, 0.1, 0.2) for b in range(B)])
    do = np.stack([env.synth.head_offsets(first + b, table.n, 1, 0.1, 0.2) for b in range(B)])
    ndet, flags = _detect_vs_c_oracle(env, "512", probs, ro, do)
    assert ndet > B * 50
    print("clustered %s B=%d: fallback rate %.4f" % (mode, B, flags[1:].mean()))
    # the sample is spatially stratified and full slices spill into a shared list: clustering must not push
    # segments onto the (slow) exact general kernels
    assert flags[1:].mean() < 0.02


def test_match_encode_bench_geometry_bit_exact(env):
    """BASELINE configs[1] exactly as bench.py launches it: ARM + ODM on B = 32 images at 512x512 in one launch
    each, every image compared with the C oracle (masks, indices, labels, matched boxes bit-exact; encodings and
    IoU bit-exact too: both sides use correctly rounded exp / log)."""
    from oracle import c_port as CP
    B, layout = 32, "512"
    table = env.otable[layout]
    _, center, labels, counts = _gt_center(env, 0, B)                        # bench.py's first match_encode batch
    gt, cb, lab, pos, idx = _arm_gpu(env, layout, center, labels, counts)
    o = CP.arm_match_encode(table, center, labels, counts, R.REFINE_POS_JAC)
    got = [flat_from_list(x, t) for x, t in ((gt, 1), (cb, 1), (lab, 1), (pos, 1), (idx, 0))]
    assert np.array_equal(got[3][..., 0], o[3]) and np.array_equal(got[4], o[4])
    assert np.array_equal(got[2][..., 0], o[2]) and bit_equal(got[1], o[1]) and bit_equal(got[0], o[0])
    ro = np.stack([env.synth.head_offsets(b, table.n) for b in range(B)])
    m = o[3] > 0                                                              # half of the positives: a head that has learnt something
    m[:, 1::2] = False
    ro[m] = o[0][m] + np.float32(0.05) * ro[m]
    ro_l = to_cuda_list(ro, table.shapes, (4,), env.dev)
    det_gt, mask, dlab, iou = env.nt.det_groundtruth(ro_l, gt, cb, lab, pos, env.anchors[layout])
    d = CP.odm_target(table, ro, o[0], o[1], o[2], o[3], R.DET_POS_JAC)
    assert np.array_equal(flat_from_list(mask, 1)[..., 0], d[1]) and np.array_equal(flat_from_list(dlab, 1)[..., 0], d[2])
    assert bit_equal(flat_from_list(iou, 0), d[3]) and bit_equal(flat_from_list(det_gt, 1), d[0])
    assert int(d[1].sum()) > B * 10


@pytest.mark.parametrize("layout,B,first", [("512", 32, 0), ("418", 5, 40), ("512", 1, 77)])
def test_target_gen_fused_bit_exact(env, layout, B, first):
    """net_tools.target_gen (ARM + ODM in one kernel, train.py:109-113 -> :147-149) against the C oracle and against
    the two-call path, every output bit for bit; B = 32 at 512x512 is bench.py's launch."""
    from oracle import c_port as CP
    table = env.otable[layout]
    _, center, labels, counts = _gt_center(env, first, B)
    o = CP.arm_match_encode(table, center, labels, counts, R.REFINE_POS_JAC)
    ro = np.stack([env.synth.head_offsets(first + b, table.n) for b in range(B)])
    m = o[3] > 0
    m[:, 1::2] = False
    ro[m] = o[0][m] + np.float32(0.05) * ro[m]
    d = CP.odm_target(table, ro, o[0], o[1], o[2], o[3], R.DET_POS_JAC)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    ro_l = to_cuda_list(ro, table.shapes, (4,), env.dev)
    arm, det = env.nt.target_gen(env.anchors[layout], t(center), t(labels), ro_l, gt_counts=t(counts), return_match_index=True)
    got = [flat_from_list(x, k) for x, k in zip(arm, (1, 1, 1, 1, 0))]
    assert np.array_equal(got[3][..., 0], o[3]) and np.array_equal(got[4], o[4]) and np.array_equal(got[2][..., 0], o[2])
    assert bit_equal(got[1], o[1]) and bit_equal(got[0], o[0])
    assert np.array_equal(flat_from_list(det[1], 1)[..., 0], d[1]) and np.array_equal(flat_from_list(det[2], 1)[..., 0], d[2])
    assert bit_equal(flat_from_list(det[3], 0), d[3]) and bit_equal(flat_from_list(det[0], 1), d[0])
    assert int(d[1].sum()) > 0
    # the two-call path gives the same bits; the lists feed the losses / torch.cat like the reference's
    a2 = env.nt.refine_groundtruth(env.anchors[layout], t(center), t(labels), env.config.refine_method.JACCARD_BIGGER, gt_counts=t(counts))
    d2 = env.nt.det_groundtruth(ro_l, a2[0], a2[1], a2[2], a2[3], env.anchors[layout])
    for x, y in zip(arm[:4], a2):
        assert torch.equal(torch.cat([v.reshape(B, -1) for v in x], 1), torch.cat([v.reshape(B, -1) for v in y], 1))
    for x, y in zip(det, d2):
        assert torch.equal(x.flat.view(torch.int32), y.flat.view(torch.int32))
    # optional outputs skipped
    arm3, det3 = env.nt.target_gen(env.anchors[layout], t(center), t(labels), ro_l, gt_counts=t(counts), need_cbboxes=False)
    assert arm3[1] is None and arm3[2] is None and torch.equal(arm3[0].flat, arm[0].flat)
    assert torch.equal(det3[0].flat.view(torch.int32), det[0].flat.view(torch.int32)) and torch.equal(det3[1].flat, det[1].flat)


def test_detect_tier2_rescues_a_too_high_cut(env):
    """The sample overestimates how many candidates lie above cut_hi (the sampled anchors carry the high scores), so tier 1
    ends up with fewer than top_k entries; the candidates between cut_lo and cut_hi (tier 2) complete the segment WITHOUT the
    general kernels, bit-exactly."""
    layout, B = "512", 2
    table = env.otable[layout]
    probs, ro, do = _detect_inputs(env, layout, 900, B, stress=False)
    sampled = np.flatnonzero((np.arange(table.n) % 19) == 9)
    others = np.flatnonzero((np.arange(table.n) % 19) != 9)
    rng = np.random.default_rng(17)
    for b in range(B):
        col = np.full(table.n, 0.01, np.float32)
        hi = rng.choice(sampled, 100, replace=False)
        col[hi] = rng.uniform(0.6, 1.0, 100).astype(np.float32)            # 34th best sample ~0.86 = cut_hi, 84th ~0.66 = cut_lo
        band = rng.choice(others, 600, replace=False)
        col[band] = rng.uniform(0.70, 0.80, 600).astype(np.float32)        # invisible to the sample, between the two cuts
        probs[b, :, 3] = col
    ndet, flags = _detect_vs_c_oracle(env, layout, probs, ro, do)
    assert ndet > 0 and flags[3].sum() == 0, "class 3 must be completed from tier 2, not by the general kernels"
