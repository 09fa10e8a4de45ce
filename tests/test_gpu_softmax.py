"""f-1: slim.softmax in front of the select stage (evaluate.py:136-137).
Parity is split in two: (1) rod_softmax follows the float32 softmax within 1e-5 relative (tolerance
parity: the reference's exp lives inside TensorFlow); (2) everything downstream is bit-exact: the
fused detected_bboxes(..., from_logits=True) equals the oracle pipeline run on rod_softmax's output,
on the sparse, dense, overflow-fallback and generic-class-count routes."""
import numpy as np
import pytest
import torch

from conftest import golden_anchors
from helpers import bit_equal, to_cuda_list
from oracle import restated as R

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5        # north star tolerance for floating point


@pytest.fixture(scope="module")
def env(cuda_device):
    from rodet_b200 import config, synth
    from rodet_b200.utils import net_tools

    class NS:
        pass
    ns = NS()
    ns.config, ns.synth, ns.nt, ns.dev = config, synth, net_tools, cuda_device
    ns.anchors = {k: golden_anchors(k) for k in ("418", "512")}
    ns.otable = {k: R.AnchorTable(v) for k, v in ns.anchors.items()}
    return ns


def _exact_softmax(z):
    z = z.astype(np.float64)
    e = np.exp(z - z.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


@pytest.mark.parametrize("shape,scale", [((1000, 11), 3.0), ((3, 7, 5, 6, 11), 3.0), ((257, 1), 1.0), ((513, 2), 10.0),
                                         ((100, 21), 5.0), ((31, 64), 2.0), ((2048, 11), 30.0), ((0, 11), 1.0)])
def test_softmax_tolerance(env, shape, scale):
    rng = np.random.default_rng(abs(hash((shape, scale))) % (2 ** 31))
    z = (rng.standard_normal(size=shape) * scale).astype(np.float32)
    x = torch.from_numpy(z).to(env.dev)
    p = env.nt.softmax(x)
    assert p.shape == x.shape and p.dtype == torch.float32
    got = p.cpu().numpy().astype(np.float64)
    if z.size == 0:
        return
    exact = _exact_softmax(z)
    assert np.all(np.abs(got - exact) <= REL_TOL * exact + 1e-30 + 1.2e-38), float(np.max(np.abs(got - exact) / (exact + 1e-30)))
    # measured accuracy is far inside the tolerance where it matters (candidates have p >= 0.01)
    big = exact >= 0.01
    if big.any():
        assert np.max(np.abs(got[big] - exact[big]) / exact[big]) < 2e-6
    assert np.allclose(got.sum(-1), 1.0, atol=1e-5)
    # float32 restatement of the reference algorithm (oracle) agrees to the same tolerance
    ref = R.softmax(z).astype(np.float64)
    assert np.all(np.abs(got - ref) <= REL_TOL * ref + 1e-30 + 1.2e-38)
    # in place, and per-layer lists
    y = x.clone()
    assert env.nt.softmax(y, out=y) is y and torch.equal(y, p)
    lst = env.nt.softmax([x, x[:1]])
    assert torch.equal(lst[0], p) and torch.equal(lst[1], p[:1])


def _inputs(env, layout, first, B, bias_class=None, bias=0.0):
    table = env.otable[layout]
    z = np.stack([env.synth.class_logits(first + b, table.n) for b in range(B)])
    if bias_class is not None:
        z[..., bias_class] += np.float32(bias)
    ro = np.stack([env.synth.head_offsets(first + b, table.n) for b in range(B)])
    do = np.stack([env.synth.head_offsets(first + b, table.n, 1) for b in range(B)])
    return table, z, ro, do


def _check_fused(env, layout, z, ro, do, sthr, nthr, topk, keep, C=11):
    table = env.otable[layout]
    logits = to_cuda_list(z, table.shapes, (C,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    probs_l = env.nt.softmax(logits)
    probs = np.concatenate([p.cpu().numpy().reshape(z.shape[0], -1, C) for p in probs_l], axis=1)
    exact = _exact_softmax(z)
    assert np.all(np.abs(probs - exact) <= REL_TOL * exact + 1e-30 + 1.2e-38)
    fs, fb, fc = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, logits, select_threshold=sthr,
                                               nms_threshold=nthr, top_k=topk, keep_top_k=keep, return_counts=True,
                                               from_logits=True)
    # (a) bit-exact against the oracle on the library's own probabilities
    o_s, o_b = R.detected_bboxes(probs, R.decode_corner(table, ro, do), sthr, nthr, None, topk, keep, num_classes=C)
    # (b) and against the unfused route softmax -> detected_bboxes
    us, ub = env.nt.decode_detected_bboxes(env.anchors[layout], ro_l, do_l, probs_l, select_threshold=sthr,
                                           nms_threshold=nthr, top_k=topk, keep_top_k=keep)
    ndet = 0
    for c in range(1, C):
        assert bit_equal(fs[c].cpu().numpy(), o_s[c]), "scores, class %d" % c
        assert bit_equal(fb[c].cpu().numpy(), o_b[c]), "boxes, class %d" % c
        assert torch.equal(fs[c], us[c]) and bit_equal(fb[c].cpu().numpy(), ub[c].cpu().numpy())
        assert np.array_equal(fc[c].cpu().numpy(), (o_s[c] != 0).sum(1))
        ndet += int((o_s[c] != 0).sum())
    return ndet, probs


@pytest.mark.parametrize("layout", ["512", "418"])
def test_fused_softmax_detect_sparse(env, layout):
    table, z, ro, do = _inputs(env, layout, 900, 3)
    ndet, _ = _check_fused(env, layout, z, ro, do, 0.3, 0.45, 400, 200)
    assert ndet > 1000


def test_fused_softmax_detect_dense_segment(env):
    """One class favoured by +6: most anchors are its candidates, so its list slices overflow and the
    second (dense) pass recomputes the same probabilities."""
    table, z, ro, do = _inputs(env, "512", 910, 2, bias_class=3, bias=6.0)
    ndet, probs = _check_fused(env, "512", z, ro, do, 0.3, 0.45, 400, 200)
    assert (probs[..., 3] >= 0.3).mean() > 0.3 and ndet > 0


def test_fused_softmax_detect_ties_take_fallback(env):
    """Constant logits: every probability is exactly 1/11, above a low threshold all of them tie, the
    candidate lists overflow and the general kernels (softmax inside their score fetch) take over."""
    table, z, ro, do = _inputs(env, "418", 920, 2)
    z[0] = np.float32(0.25)
    ndet, probs = _check_fused(env, "418", z, ro, do, 0.05, 0.45, 400, 200)
    assert np.all(probs[0] == probs[0, 0, 0]) and ndet > 0


def test_fused_softmax_detect_script_defaults_and_low_threshold(env):
    table, z, ro, do = _inputs(env, "418", 930, 1)
    _check_fused(env, "418", z, ro, do, 0.1, 0.4, 400, 200)          # predict.py:136-137
    _check_fused(env, "418", z, ro, do, 0.02, 0.5, 800, 200)         # signature default top_k, many candidates


def test_fused_softmax_generic_class_count(env, monkeypatch):
    """Class depth != 11: no fused scan; the general path applies the softmax when it fetches scores."""
    layout, B, C = "418", 2, 6
    table = env.otable[layout]
    rng = np.random.default_rng(78)
    z = (rng.standard_normal(size=(B, table.n, C)) * 3.0).astype(np.float32)
    z[..., 0] += np.float32(2.0)
    ro = np.stack([env.synth.head_offsets(940 + b, table.n) for b in range(B)])
    do = np.stack([env.synth.head_offsets(940 + b, table.n, 1) for b in range(B)])
    monkeypatch.setattr(env.config, "total_obj_n", C)
    ndet, _ = _check_fused(env, layout, z, ro, do, 0.3, 0.45, 400, 200, C=C)
    assert ndet > 0


def test_logits_with_localisations_entry(env):
    """detected_bboxes(logits, localisations, from_logits=True): the reference call order of
    evaluate.py:136-151 with only the softmax folded in."""
    from rodet_b200.utils import common_tools as ct
    table, z, ro, do = _inputs(env, "418", 950, 2)
    logits = to_cuda_list(z, table.shapes, (11,), env.dev)
    ro_l, do_l = to_cuda_list(ro, table.shapes, (4,), env.dev), to_cuda_list(do, table.shapes, (4,), env.dev)
    locs = [ct.centerBboxes_2_cornerBboxes(env.nt.decode_locations_one_layer(a, r + d))
            for a, r, d in zip(env.anchors["418"], ro_l, do_l)]
    fs, fb = env.nt.detected_bboxes(logits, locs, select_threshold=0.3, nms_threshold=0.45, top_k=400, keep_top_k=200,
                                    from_logits=True)
    us, ub = env.nt.detected_bboxes(env.nt.softmax(logits), locs, select_threshold=0.3, nms_threshold=0.45, top_k=400,
                                    keep_top_k=200)
    for c in range(1, 11):
        assert torch.equal(fs[c], us[c]) and torch.equal(fb[c], ub[c])
    assert sum(int((fs[c] != 0).sum()) for c in fs) > 0
