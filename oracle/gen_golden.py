"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (Tier A).

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is
mounted:      python -m oracle.gen_golden
The reference sources (utils/net_tools.py, utils/common_tools.py,
utils/tf_extended/*.py) are imported as they lie and run over oracle/tf_shim, an
eager NumPy stand-in for the TensorFlow-1 symbols they call.  The fixtures travel
to the GPU box, where /root/reference does not exist.

Every fixture stores its inputs in full (small cases) or the generator seed plus
an input checksum (full-size cases), and the reference's outputs.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader                    # noqa: E402
from oracle.tf_shim import to_numpy               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FEATS = {
    "418": ((418, 418), [(53, 53), (27, 27), (14, 14), (7, 7), (4, 4), (2, 2)]),
    "512": ((512, 512), [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]),
    "tiny": ((96, 96), [(6, 5), (3, 3), (2, 2), (1, 2), (1, 1), (1, 1)]),
}


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    return h.hexdigest()


def flat(per_layer, tail):
    outs = []
    for t in per_layer:
        t = np.asarray(t)
        lead = t.ndim - 3 - tail
        outs.append(t.reshape(t.shape[:lead] + (-1,) + t.shape[t.ndim - tail:]))
    return np.concatenate(outs, axis=outs[0].ndim - 1 - tail)


def split_layers(flat_arr, shapes, tail):
    """[B,N,*tail] -> list of [B,fh,fw,A,*tail]."""
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        sl = flat_arr[:, off:off + n]
        out.append(sl.reshape((flat_arr.shape[0], fh, fw, a) + flat_arr.shape[2:]))
        off += n
    return out


def layer_shapes(anchors):
    return [(a[0].shape[0], a[0].shape[1], a[2].shape[0]) for a in anchors]


def rand_gt(rng, g, dup=False):
    c = rng.uniform(0.1, 0.9, size=(g, 2))
    hw = np.exp(rng.uniform(np.log(0.03), np.log(0.6), size=(g, 2)))
    cr = np.clip(np.concatenate([c - hw / 2, c + hw / 2], 1), 0, 1).astype(np.float32)
    if dup and g >= 3:
        cr[g - 1] = cr[0]            # exact duplicate: argmax must report the lower index
        cr[g // 2] = cr[1]
    labels = rng.integers(1, 11, size=g).astype(np.int64)
    return cr, labels


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    tf, nt, ct, tfe, cfg = ref.tf, ref.net_tools, ref.common_tools, ref.tfe, ref.config

    # ------------------------------------------------------------------ anchors
    store = {}
    for name, (img, feats) in FEATS.items():
        anc = ref_loader.reference_anchors(ref, img, feats)
        for l, (y, x, h, w) in enumerate(anc):
            store["%s_l%d_y" % (name, l)] = y
            store["%s_l%d_x" % (name, l)] = x
            store["%s_l%d_h" % (name, l)] = h
            store["%s_l%d_w" % (name, l)] = w
        store["%s_img" % name] = np.asarray(img)
        store["%s_feats" % name] = np.asarray(feats)
        sizes = nt.init_anchor(len(feats))
        store["%s_sizes_px" % name] = np.concatenate(list(sizes.values()))
    store["n_anchor_each_layer"] = np.asarray(nt.n_anchor_each_layer("mobilenet_v2"))
    np.savez_compressed(os.path.join(OUT, "anchors.npz"), **store)

    # ------------------------------------------------------------------ box helpers
    rng = np.random.default_rng(11)
    cr = np.sort(rng.uniform(-0.2, 1.2, size=(3, 7, 2, 2)).astype(np.float32), axis=2).reshape(3, 7, 4)
    ce = to_numpy(ct.cornerBboxes_2_centerBboxes(tf.constant(cr)))
    back = to_numpy(ct.centerBboxes_2_cornerBboxes(tf.constant(ce)))
    np.savez_compressed(os.path.join(OUT, "box_format.npz"), corner=cr, center=ce, corner_back=back)

    # ------------------------------------------------------------------ ARM / ODM / encode / decode / jaccard
    cases = [  # name, layout, G, dup, seed
        ("tiny_g1", "tiny", 1, False, 1), ("tiny_g2", "tiny", 2, False, 2),
        ("tiny_g5_dup", "tiny", 5, True, 3), ("tiny_g8", "tiny", 8, False, 4),
        ("tiny_g8_dup", "tiny", 8, True, 5),
        ("r418_g3", "418", 3, False, 6), ("r418_g17_dup", "418", 17, True, 7),
        ("r512_g100", "512", 100, False, 8), ("r512_g41_dup", "512", 41, True, 9),
    ]
    for name, layout, g, dup, seed in cases:
        img, feats = FEATS[layout]
        anc = ref_loader.reference_anchors(ref, img, feats)
        shapes = layer_shapes(anc)
        rng = np.random.default_rng(seed)
        corner, labels = rand_gt(rng, g, dup)
        if layout == "tiny":
            # derive the GT from randomly chosen anchors (jittered) so that the tiny
            # layouts have positives at several IoU levels around the thresholds
            for i in range(g):
                l = int(rng.integers(1, len(anc)))
                y, x, h, w = anc[l]
                fy, fx, a = (int(rng.integers(0, d)) for d in (y.shape[0], y.shape[1], h.shape[0]))
                sc_h, sc_w = rng.uniform(0.6, 1.5, size=2)
                dy, dx = rng.uniform(-0.25, 0.25, size=2) * (h[a], w[a])
                corner[i] = np.clip(np.array([y[fy, fx, 0] + dy - sc_h * h[a] / 2, x[fy, fx, 0] + dx - sc_w * w[a] / 2,
                                              y[fy, fx, 0] + dy + sc_h * h[a] / 2, x[fy, fx, 0] + dx + sc_w * w[a] / 2]),
                                    0, 1).astype(np.float32)
            if dup and g >= 3:
                corner[g - 1] = corner[0]
                corner[g // 2] = corner[1]
        if name == "tiny_g8":
            # a GT exactly equal to one anchor's corner box (IoU ~ 1) and a tiny far box
            y, x, h, w = anc[1]
            corner[0] = np.array([y[1, 1, 0] - h[3] / 2, x[1, 1, 0] - w[3] / 2,
                                  y[1, 1, 0] + h[3] / 2, x[1, 1, 0] + w[3] / 2], np.float32)
            corner[1] = np.array([0.001, 0.001, 0.004, 0.003], np.float32)
        center = to_numpy(ct.cornerBboxes_2_centerBboxes(tf.constant(corner)))
        out = {"layout": layout, "corner": corner, "center": center, "labels": labels}
        for method in ("JACCARD_BIGGER", "NEAREST_NEIGHBOR"):
            if method == "NEAREST_NEIGHBOR" and layout != "tiny":
                continue
            r = nt.refine_groundtruth(anc, tf.constant(center), tf.constant(labels, dtype=np.int64),
                                      getattr(cfg.refine_method, method))
            gt, cb, lab, pos = [to_numpy(v) for v in r]
            tag = "jb" if method == "JACCARD_BIGGER" else "nn"
            out[tag + "_gt"] = flat(gt, 1)
            out[tag + "_cb"] = flat(cb, 1)
            out[tag + "_labels"] = flat(lab, 1)[..., 0]
            out[tag + "_pos"] = flat(pos, 1)[..., 0]
        # the reference does not return maxJacIndex; recompute it with the reference's
        # own jaccard() so the fixture also pins the argmax (ties -> first)
        idx_layers = []
        for (y, x, h, w) in anc:
            ca = tf.stack([np.float32(y - h / 2.), np.float32(x - w / 2.),
                           np.float32(y + h / 2.), np.float32(x + w / 2.)], axis=-1)
            jac = np.stack([to_numpy(nt.jaccard(ca, ct.centerBboxes_2_cornerBboxes(tf.constant(center)[i])))
                            for i in range(g)])
            idx_layers.append(np.argmax(jac, axis=0).astype(np.int32).reshape(-1))
        out["jb_idx"] = np.concatenate(idx_layers)

        # ODM on the ARM result, B=2 random refine_out (utils/net_tools.py:431-475)
        n = out["jb_pos"].shape[0]
        B = 2
        ro = (rng.standard_normal(size=(B, n, 4)) * np.array([0.1, 0.1, 0.2, 0.2])).astype(np.float32)
        # make some refined anchors land close to the GT so that ODM positives exist
        ro = np.where(out["jb_pos"][None, :, None] > 0,
                      (out["jb_gt"][None] + rng.uniform(0, 2.5, size=(B, n, 1)) * ro).astype(np.float32), ro)
        og = np.broadcast_to(out["jb_gt"], (B, n, 4)).copy()
        cbb = np.broadcast_to(out["jb_cb"], (B, n, 4)).copy()
        lb = np.broadcast_to(out["jb_labels"], (B, n)).copy()
        pm = np.broadcast_to(out["jb_pos"], (B, n)).copy()
        r = nt.det_groundtruth([tf.constant(v) for v in split_layers(ro, shapes, 1)],
                               [tf.constant(v) for v in split_layers(og, shapes, 1)],
                               [tf.constant(v) for v in split_layers(cbb, shapes, 1)],
                               [tf.constant(v[..., None]) for v in split_layers(lb, shapes, 0)],
                               [tf.constant(v[..., None]) for v in split_layers(pm, shapes, 0)],
                               anc)
        det_gt, mask, det_lab, iou = [to_numpy(v) for v in r]
        out["odm_refine_out"] = ro
        out["odm_det_gt"] = flat(det_gt, 1)
        out["odm_mask"] = flat(mask, 1)[..., 0]
        out["odm_labels"] = flat(det_lab, 1)[..., 0]
        out["odm_iou"] = flat(iou, 0)

        # decode call site (evaluate.py:139-143): c2c(decode(anchors, refine+det))
        do = (rng.standard_normal(size=(B, n, 4)) * np.array([0.1, 0.1, 0.2, 0.2])).astype(np.float32)
        locs = []
        for a_l, r_l, d_l in zip(anc, split_layers(ro, shapes, 1), split_layers(do, shapes, 1)):
            c = nt.decode_locations_one_layer(a_l, tf.constant(r_l) + tf.constant(d_l))
            locs.append(to_numpy(ct.centerBboxes_2_cornerBboxes(c)))
        out["dec_det_out"] = do
        out["dec_corner"] = flat(locs, 1)
        if layout == "tiny":
            enc = [to_numpy(nt.encode_locations_one_layer(a_l, tf.constant(center)[0])) for a_l in anc]
            out["enc_gt0"] = flat(enc, 1)
        if layout != "tiny":
            # keep the full-size fixtures small: sparse storage of the dense outputs
            for k in ("jb_gt", "jb_cb"):
                out[k + "_sparse"] = out[k][out["jb_pos"] > 0]
                out[k + "_digest"] = digest(out.pop(k))
            out["odm_det_gt_sparse"] = out["odm_det_gt"][out["odm_mask"] > 0]
            out["odm_det_gt_digest"] = digest(out.pop("odm_det_gt"))
            out["odm_iou"] = out["odm_iou"].astype(np.float32)
            out["odm_refine_out_digest"] = digest(ro); out.pop("odm_refine_out")
            out["dec_det_out_digest"] = digest(do); out.pop("dec_det_out")
            out["dec_corner"] = out["dec_corner"][:, ::37]     # strided sample
            out["seed"] = seed
        np.savez_compressed(os.path.join(OUT, "targets_%s.npz" % name), **out)
        print("targets", name, "pos", int(out["jb_pos"].sum()), "odm pos", int(out["odm_mask"].sum()))

    # ------------------------------------------------------------------ select / sort / NMS / detected_bboxes
    det_cases = [  # name, layout, B, select_thr, nms_thr, top_k, keep, mode, seed
        ("tiny_a", "tiny", 2, 0.3, 0.45, 40, 12, "normal", 21),
        ("tiny_none", "tiny", 2, None, 0.5, 30, 10, "normal", 22),      # None -> 0.0: all pass
        ("tiny_dupscores", "tiny", 3, 0.2, 0.4, 50, 20, "quantised", 23),
        ("tiny_stress", "tiny", 2, 0.3, 0.45, 60, 25, "stress", 24),
        ("tiny_clip", "tiny", 1, 0.1, 0.4, 40, 15, "normal", 25),
        ("r418_eval", "418", 1, 0.3, 0.4, 400, 200, "normal", 26),       # evaluate.py:58-65
        ("r418_stress", "418", 1, 0.3, 0.45, 400, 200, "stress", 27),
    ]
    for name, layout, B, sthr, nthr, topk, keep, mode, seed in det_cases:
        img, feats = FEATS[layout]
        anc = ref_loader.reference_anchors(ref, img, feats)
        shapes = layer_shapes(anc)
        n = sum(a * b * c for a, b, c in shapes)
        rng = np.random.default_rng(seed)
        if mode == "stress":
            probs = rng.uniform(sthr, 1.0, size=(B, n, 11)).astype(np.float32)
            sig = np.array([0.05, 0.05, 0.05, 0.05])
        else:
            z = (rng.standard_normal(size=(B, n, 11)) * 3.0).astype(np.float32)
            z[..., 0] += np.float32(1.0 if layout == "tiny" else 4.0)
            z -= z.max(-1, keepdims=True)
            e = np.exp(z)
            probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
            sig = np.array([0.1, 0.1, 0.2, 0.2])
            if mode == "quantised":       # many exactly equal scores -> tie rules matter
                probs = (np.round(probs * 8) / 8).astype(np.float32)
        ro = (rng.standard_normal(size=(B, n, 4)) * sig).astype(np.float32)
        do = (rng.standard_normal(size=(B, n, 4)) * sig).astype(np.float32)
        locs = []
        for a_l, r_l, d_l in zip(anc, split_layers(ro, shapes, 1), split_layers(do, shapes, 1)):
            c = nt.decode_locations_one_layer(a_l, tf.constant(r_l) + tf.constant(d_l))
            locs.append(ct.centerBboxes_2_cornerBboxes(c))
        preds = [tf.constant(v) for v in split_layers(probs, shapes, 1)]
        clip = np.array([0.1, 0.1, 0.8, 0.9], np.float32) if name == "tiny_clip" else None
        rs, rb = nt.detected_bboxes(preds, locs, select_threshold=sthr, nms_threshold=nthr,
                                    clipping_bbox=None if clip is None else tf.constant(clip),
                                    top_k=topk, keep_top_k=keep)
        out = {"layout": layout, "B": B, "select_threshold": np.float32(-1 if sthr is None else sthr),
               "select_none": sthr is None, "nms_threshold": np.float32(nthr), "top_k": topk,
               "keep_top_k": keep, "seed": seed, "mode": mode,
               "boxes_corner": flat(to_numpy(locs), 1)}
        if clip is not None:
            out["clip"] = clip
        for c in rs:
            out["scores_c%d" % c] = to_numpy(rs[c])
            out["bboxes_c%d" % c] = to_numpy(rb[c])
        if layout == "tiny":
            out.update(probs=probs, refine_out=ro, det_out=do)
            # stage outputs too: select -> sort (before NMS)
            d_s, d_b = nt.bboxes_select_all_layers(preds, locs, select_threshold=sthr,
                                                   num_classes=cfg.total_obj_n)
            s_s, s_b = tfe.bboxes_sort(d_s, d_b, top_k=topk)
            for c in d_s:
                out["sel_scores_c%d" % c] = to_numpy(d_s[c])
                out["sel_bboxes_c%d" % c] = to_numpy(d_b[c])
                out["sort_scores_c%d" % c] = to_numpy(s_s[c])
                out["sort_bboxes_c%d" % c] = to_numpy(s_b[c])
        else:
            out.update(probs_digest=digest(probs), refine_out_digest=digest(ro), det_out_digest=digest(do))
            out.pop("boxes_corner")
        np.savez_compressed(os.path.join(OUT, "detect_%s.npz" % name), **out)
        print("detect", name, "nonzero dets", sum(int((to_numpy(rs[c]) > 0).sum()) for c in rs))

    # ------------------------------------------------------------------ standalone tfe ops
    rng = np.random.default_rng(31)
    boxes = np.sort(rng.uniform(0, 1, size=(60, 2, 2)).astype(np.float32), axis=1).reshape(60, 4)
    boxes[7] = boxes[3]; boxes[11] = np.array([0.5, 0.5, 0.5, 0.7], np.float32)   # dup + zero-area
    refb = np.array([0.2, 0.1, 0.7, 0.9], np.float32)
    sc = (np.round(rng.uniform(0, 1, size=60) * 16) / 16).astype(np.float32)
    o = {"boxes": boxes, "ref": refb, "scores": sc}
    o["jaccard"] = to_numpy(tfe.bboxes_jaccard(tf.constant(refb), tf.constant(boxes)))
    o["intersection"] = to_numpy(tfe.bboxes_intersection(tf.constant(refb), tf.constant(boxes)))
    o["resize"] = to_numpy(tfe.bboxes_resize(tf.constant(refb), tf.constant(boxes)))
    o["clip"] = to_numpy(tfe.bboxes_clip(tf.constant(refb), tf.constant(boxes)))
    s, b = tfe.bboxes_nms(tf.constant(sc), tf.constant(boxes), nms_threshold=0.3, keep_top_k=25)
    o["nms_scores"], o["nms_bboxes"] = to_numpy(s), to_numpy(b)
    s, b = tfe.bboxes_sort(tf.constant(sc[None]), tf.constant(boxes[None]), top_k=20)
    o["sort_scores"], o["sort_bboxes"] = to_numpy(s), to_numpy(b)
    o["pad"] = to_numpy(tfe.pad_axis(tf.constant(boxes[:5]), 0, 9, axis=0))
    np.savez_compressed(os.path.join(OUT, "tfe_ops.npz"), **o)
    print("done ->", OUT)


if __name__ == "__main__":
    main()
