"""f-4: the GT-box chain of the training input pipeline (resize to the crop, overlap filter, flip, clamp)
fused into one batched kernel, against the reference-generated fixture and the oracle on BASELINE-shaped GT,
then straight into refine_groundtruth."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_anchors
from helpers import bit_equal
from oracle import restated as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda_device):
    import rodet_b200.utils.tf_extended as tfe
    from rodet_b200 import config, synth
    from rodet_b200.utils import common_tools, data_pileline_tools, net_tools

    class NS:
        pass
    ns = NS()
    ns.tfe, ns.dp, ns.nt, ns.ct, ns.synth, ns.config, ns.dev = tfe, data_pileline_tools, net_tools, common_tools, synth, config, cuda_device
    return ns


@pytest.mark.parametrize("neg", [False, True])
@pytest.mark.parametrize("i64", [True, False])
def test_fixture_batched(env, neg, i64):
    z = golden("gt_boxes.npz")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    labels = z["labels"] if i64 else z["labels"].astype(np.int32)
    ol, ob, oc = env.dp.process_raw_gt_train(d(labels), d(z["boxes"]), d(z["counts"]), d(z["crops"]), d(z["mirror"]),
                                             assign_negative=neg)
    assert ol.dtype == (torch.int64 if i64 else torch.int32) and oc.dtype == torch.int32
    ol, ob, oc = ol.cpu().numpy(), ob.cpu().numpy(), oc.cpu().numpy()
    for b in range(int(z["B"])):
        el, eb = z["labels_%d_%d" % (b, int(neg))], z["bboxes_%d_%d" % (b, int(neg))]
        k = el.size
        assert oc[b] == k, (b, oc[b], k)
        assert np.array_equal(ol[b, :k], el) and bit_equal(ob[b, :k], eb), b
        assert not ol[b, k:].any() and not ob[b, k:].any()              # zero padding behind the kept boxes


def test_filter_overlap_single_image(env):
    z = golden("gt_boxes.npz")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    b = 0
    n = int(z["counts"][b])
    resized = env.tfe.bboxes_resize(d(z["crops"][b]), d(z["boxes"][b, :n]))
    lab, bx = env.tfe.bboxes_filter_overlap(d(z["labels"][b, :n]), resized, threshold=0.3)
    el, eb = R.bboxes_filter_overlap(z["labels"][b, :n], R.bboxes_resize(z["crops"][b], z["boxes"][b, :n]), 0.3)
    assert np.array_equal(lab.cpu().numpy(), el) and bit_equal(bx.cpu().numpy(), eb) and 0 < el.size < n
    lab2, bx2 = env.tfe.bboxes_filter_overlap(d(z["labels"][b, :n]), resized, threshold=0.3, assign_negative=True)
    el2, eb2 = R.bboxes_filter_overlap(z["labels"][b, :n], R.bboxes_resize(z["crops"][b], z["boxes"][b, :n]), 0.3, True)
    assert np.array_equal(lab2.cpu().numpy(), el2) and bit_equal(bx2.cpu().numpy(), eb2) and (el2 < 0).any()
    # empty input
    e_l, e_b = env.tfe.bboxes_filter_overlap(torch.zeros(0, dtype=torch.int64, device=env.dev), torch.zeros(0, 4, device=env.dev))
    assert e_l.numel() == 0 and e_b.shape == (0, 4)


def test_random_batch_into_arm(env):
    """BASELINE-shaped GT (up to 100 boxes, B = 32) through the fused chain, then corner -> centre -> ARM
    with the new counts: every stage against the oracle."""
    B = 32
    corner, labels, counts = env.synth.gt_batch(5000, B)
    rng = np.random.default_rng(8)
    crops = np.stack([np.concatenate([rng.uniform(0, 0.35, 2), rng.uniform(0.6, 1.0, 2)]) for _ in range(B)]).astype(np.float32)
    mirror = rng.uniform(size=B) < 0.5
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    ol, ob, oc = env.dp.process_raw_gt_train(d(labels), d(corner), d(counts), d(crops), d(mirror))
    g_l, g_b, g_c = ol.cpu().numpy(), ob.cpu().numpy(), oc.cpu().numpy()
    dropped = 0
    for b in range(B):
        el, eb = R.gt_boxes_train(labels[b, :counts[b]], corner[b, :counts[b]], crops[b], bool(mirror[b]))
        assert g_c[b] == el.size and np.array_equal(g_l[b, :el.size], el) and bit_equal(g_b[b, :el.size], eb), b
        assert eb.min(initial=0) >= 0 and eb.max(initial=0) <= 1
        dropped += counts[b] - el.size
    assert dropped > 0
    anchors = golden_anchors("418")
    table = R.AnchorTable(anchors)
    center = env.ct.cornerBboxes_2_centerBboxes(ob)
    gt, cb, lab, pos, idx = env.nt.refine_groundtruth(anchors, center, ol, env.config.refine_method.JACCARD_BIGGER, gt_counts=oc,
                                                      return_match_index=True)
    gpos = np.concatenate([p.cpu().numpy().reshape(B, -1) for p in pos], axis=1)
    gidx = np.concatenate([p.cpu().numpy().reshape(B, -1) for p in idx], axis=1)
    for b in (0, 7, 31):
        k = int(g_c[b])
        if k == 0:
            continue
        o = R.arm_match_encode(table, R.corner_to_center(g_b[b, :k]), g_l[b, :k])
        assert np.array_equal(gpos[b], o[3]) and np.array_equal(gidx[b], o[4])


def test_tfrecord_ground_truth_batches_on_device(env):
    """f-4: records written by the reference converter -> native reader -> device-resident ragged arrays -> padded batches
    assembled by rod_gt_gather -> straight into the box chain and ARM matching."""
    import os
    from conftest import GOLDEN
    from rodet_b200.dataset.pascalvoc_common import read_ground_truth
    z = golden("voc_gt_expected.npz")
    gt = read_ground_truth(os.path.join(GOLDEN, "voc_gt_000.tfrecord"))
    dgt = gt.to(env.dev)
    idx = [7, 0, 5, 11, 3]                                              # record 5 has no objects
    bboxes, labels, diff, counts = dgt.batch(idx)
    G = int(np.diff(z["offsets"]).max())
    assert bboxes.shape == (5, G, 4) and labels.shape == (5, G) and counts.dtype == torch.int32
    b_np, l_np, c_np = bboxes.cpu().numpy(), labels.cpu().numpy(), counts.cpu().numpy()
    for row, r in enumerate(idx):
        o0, o1 = int(z["offsets"][r]), int(z["offsets"][r + 1])
        g = o1 - o0
        assert c_np[row] == g
        exp = np.stack([z["ymin"][o0:o1], z["xmin"][o0:o1], z["ymax"][o0:o1], z["xmax"][o0:o1]], 1)
        assert bit_equal(b_np[row, :g], exp) and np.array_equal(l_np[row, :g], z["label"][o0:o1])
        assert not b_np[row, g:].any() and not l_np[row, g:].any()
    assert not diff.any()
    # clipped to max_gt, and the default index range
    b2, l2, _, c2 = dgt.batch(max_gt=4, batch=3)
    assert b2.shape == (3, 4, 4) and c2.tolist() == [min(4, int(z["offsets"][r + 1] - z["offsets"][r])) for r in range(3)]
    with pytest.raises(IndexError):
        dgt.batch([12])
    # device-resident indices are not read back: an invalid one gives a zero row and counts = -1
    b3, l3, _, c3 = dgt.batch(torch.tensor([3, 12, -1, 7], device=env.dev))
    assert c3.tolist() == [int(z["offsets"][4] - z["offsets"][3]), -1, -1, int(z["offsets"][8] - z["offsets"][7])]
    assert not b3[1:3].any() and not l3[1:3].any() and torch.equal(b3[3], bboxes[0])
    with pytest.raises(ValueError):
        gt.batch([0])                                                   # host arrays: no CPU path
    # into the pipeline: clamp chain -> centre form -> ARM (images without objects give all-zero targets)
    ol, ob, oc = env.dp.process_raw_gt_train(labels, bboxes, counts)
    center = env.ct.cornerBboxes_2_centerBboxes(ob)
    anchors = golden_anchors("418")
    out = env.nt.refine_groundtruth(anchors, center, ol, env.config.refine_method.JACCARD_BIGGER, gt_counts=oc)
    pos = np.concatenate([p.cpu().numpy().reshape(5, -1) for p in out[3]], axis=1)
    assert pos[2].sum() == 0 and pos[0].sum() > 0
    table = R.AnchorTable(anchors)
    k = int(oc[0])
    o = R.arm_match_encode(table, R.corner_to_center(ob[0, :k].cpu().numpy()), ol[0, :k].cpu().numpy())
    assert np.array_equal(pos[0], o[3])
