"""f-2 evaluation metrics on the GPU (evaluate.py:162-197): streaming TP/FP accumulation, precision /
recall and VOC07 / VOC12 average precision, against the fixture written by the unmodified reference
(tests/golden/eval_metrics.npz) and against oracle/restated on streamed random batches."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import restated as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tfe(cuda_device):
    import rodet_b200.utils.tf_extended as m
    m.dev = cuda_device
    return m


def _ap_close(a, b):
    return abs(float(a) - float(b)) <= 1e-12 * max(1.0, abs(float(b)))


def test_precision_recall_and_ap_vs_reference_fixture(tfe):
    z = golden("eval_metrics.npz")
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(tfe.dev)
    for i in range(int(z["n_cases"])):
        s, tp, fp, ngb = z["scores_%d" % i], z["tp_%d" % i], z["fp_%d" % i], int(z["ngb_%d" % i])
        p, r = tfe.precision_recall(torch.tensor(ngb, device=tfe.dev), s.size, d(tp), d(fp), d(s))
        assert p.dtype == torch.float64 and r.dtype == torch.float64
        assert np.array_equal(p.cpu().numpy(), z["precision_%d" % i]), i      # bit-exact float64
        assert np.array_equal(r.cpu().numpy(), z["recall_%d" % i]), i
        assert float(tfe.average_precision_voc07(p, r)) == float(z["voc07_%d" % i]), i   # summed in the reference's order
        assert _ap_close(tfe.average_precision_voc12(p, r), z["voc12_%d" % i]), i        # reduce_sum order is TF's
        p32, _ = tfe.precision_recall(ngb, s.size, d(tp), d(fp), d(s), dtype=torch.float32)
        assert p32.dtype == torch.float32
    x = d(z["cummax_in"])
    assert np.array_equal(tfe.cummax(x).cpu().numpy(), z["cummax_fwd"])
    assert np.array_equal(tfe.cummax(x, reverse=True).cpu().numpy(), z["cummax_rev"])


@pytest.mark.parametrize("remove_zero", [True, False])
def test_streaming_accumulation_dict(tfe, remove_zero):
    """Several batches of per-class [B, keep] arrays through streaming_tp_fp_arrays (dict form), growing the
    device buffers, then precision_recall + AP per class: arrays bit-equal to the restated local variables."""
    tfe.reset_local_variables()
    rng = np.random.default_rng(3)
    classes, B, keep = [1, 2, 5], 16, 200
    ref = {c: R.StreamingTpFp() for c in classes}
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(tfe.dev)
    vals = None
    for it in range(6):                                     # 6 x 3200 slots: forces two reallocations from 4096
        num_gb, tp, fp, sc = {}, {}, {}, {}
        for c in classes:
            s = np.sort(rng.uniform(0, 1, size=(B, keep)).astype(np.float32), axis=1)[:, ::-1].copy()
            s = (np.round(s * 64) / 64).astype(np.float32)  # ties across images
            s[:, rng.integers(100, keep):] = 0              # pad_axis zero padding of the NMS output
            t = (rng.uniform(size=(B, keep)) < 0.4) & (s > 0)
            f = (rng.uniform(size=(B, keep)) < 0.7) & ~t
            n = rng.integers(0, 9, size=B).astype(np.int64)
            ref[c].update(n, t, f, s, remove_zero)
            num_gb[c], tp[c], fp[c], sc[c] = d(n), d(t), d(f), d(s)
        vals, upd = tfe.streaming_tp_fp_arrays(num_gb, tp, fp, sc, remove_zero_scores=remove_zero)
        assert set(vals.keys()) == set(classes) and set(upd.keys()) == set(classes)
    for c in classes:
        no, nd, t, f, s = vals[c]
        ro, rd, rt, rf, rs = ref[c].value()
        assert int(no) == int(ro) and int(nd) == int(rd) and nd.dtype == torch.int32
        assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(t.cpu().numpy(), rt) and np.array_equal(f.cpu().numpy(), rf)
        p, r = tfe.precision_recall(*vals[c])
        op, orc = R.precision_recall(ro, rd, rt, rf, rs)
        assert np.array_equal(p.cpu().numpy(), op) and np.array_equal(r.cpu().numpy(), orc)
        assert float(tfe.average_precision_voc07(p, r)) == float(R.average_precision_voc07(op, orc))
        assert _ap_close(tfe.average_precision_voc12(p, r), R.average_precision_voc12(op, orc))
    # dict form of precision_recall (evaluate.py:178 loops over classes; the reference accepts dicts too)
    dp, dr = tfe.precision_recall({c: vals[c][0] for c in classes}, {c: vals[c][1] for c in classes},
                                  {c: vals[c][2] for c in classes}, {c: vals[c][3] for c in classes},
                                  {c: vals[c][4] for c in classes})
    assert set(dp.keys()) == set(classes) and dp[1].numel() == int(vals[1][1])
    tfe.reset_local_variables()


def test_streaming_single_tensor_and_empty(tfe):
    tfe.reset_local_variables()
    dev = tfe.dev
    s = torch.tensor([[0.9, 0.5, 0.0], [0.00005, 0.3, 0.2]], device=dev)
    tp = torch.tensor([[1, 0, 0], [1, 0, 1]], dtype=torch.bool, device=dev)
    fp = torch.tensor([[0, 1, 0], [0, 0, 0]], dtype=torch.bool, device=dev)
    (no, nd, t, f, sc), _ = tfe.streaming_tp_fp_arrays(torch.tensor([2, 1], device=dev), tp, fp, s, name="one")
    assert int(no) == 3 and int(nd) == 3
    assert sc.tolist() == pytest.approx([0.9, 0.5, 0.2]) and t.tolist() == [True, False, True] and f.tolist() == [False, True, False]
    # nothing survives the filter: empty arrays, AP 0
    (no, nd, t, f, sc), _ = tfe.streaming_tp_fp_arrays(torch.tensor([4], device=dev), torch.zeros(1, 5, dtype=torch.bool, device=dev),
                                                       torch.zeros(1, 5, dtype=torch.bool, device=dev), torch.rand(1, 5, device=dev),
                                                       name="empty")
    assert int(no) == 4 and int(nd) == 0 and sc.numel() == 0
    p, r = tfe.precision_recall(no, nd, t, f, sc)
    assert p.numel() == 0 and float(tfe.average_precision_voc07(p, r)) == 0.0 and float(tfe.average_precision_voc12(p, r)) == 0.0
    with pytest.raises(ValueError, match="top_k"):
        tfe.precision_recall(1, 9, tp.reshape(-1), fp.reshape(-1), s.reshape(-1))
    tfe.reset_local_variables()


def test_large_scan_many_tiles(tfe):
    """1.3 M detections: > 300 scan tiles in rod_precision_recall and > 1200 reverse tiles in the AP kernel."""
    rng = np.random.default_rng(11)
    n, ngb = 1_300_007, 400_000
    s = rng.uniform(0.01, 1, size=n).astype(np.float32)
    tp = rng.uniform(size=n) < 0.3 * (0.5 + s)
    fp = ~tp
    d = lambda a: torch.from_numpy(a).to(tfe.dev)
    p, r = tfe.precision_recall(ngb, n, d(tp), d(fp), d(s))
    op, orc = R.precision_recall(ngb, n, tp, fp, s)
    assert np.array_equal(p.cpu().numpy(), op) and np.array_equal(r.cpu().numpy(), orc)
    assert float(tfe.average_precision_voc07(p, r)) == float(R.average_precision_voc07(op, orc))
    assert _ap_close(tfe.average_precision_voc12(p, r), R.average_precision_voc12(op, orc))


def test_end_to_end_eval_chain(tfe):
    """detections -> bboxes_matching_batch -> streaming_tp_fp_arrays -> precision_recall -> AP, the call
    sequence of evaluate.py:153-197, against the oracle chained the same way."""
    tfe.reset_local_variables()
    z = golden("eval_matching.npz")
    classes = [int(c) for c in z["classes"]]
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(tfe.dev)
    rs = {c: d(z["scores_c%d" % c]) for c in classes}
    rb = {c: d(z["bboxes_c%d" % c]) for c in classes}
    n, tp, fp, sc = tfe.bboxes_matching_batch(classes, rs, rb, d(z["glabels"]), d(z["gbboxes"]), d(z["gdifficults"]), 0.5)
    vals, _ = tfe.streaming_tp_fp_arrays(n, tp, fp, sc)
    aps = []
    for c in classes:
        st = R.StreamingTpFp()
        st.update(z["n_c%d" % c], z["tp_c%d" % c], z["fp_c%d" % c], z["scores_c%d" % c])
        op, orc = R.precision_recall(*st.value())
        p, r = tfe.precision_recall(*vals[c])
        assert np.array_equal(p.cpu().numpy(), op) and np.array_equal(r.cpu().numpy(), orc)
        v = tfe.average_precision_voc07(p, r)
        assert float(v) == float(R.average_precision_voc07(op, orc))
        aps.append(v)
    mAP = torch.stack(aps).sum() / len(aps)                # evaluate.py:184-185  tf.add_n(aps) / len(aps)
    assert 0.0 <= float(mAP) <= 1.0
    tfe.reset_local_variables()


@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 5000, 70001])
def test_score_sort_matches_top_k_order(cuda_device, n):
    """rod_sort_scores_desc (the library's own sort behind precision_recall): descending score, equal scores in
    index order = tf.nn.top_k(scores, k, sorted=True) (utils/tf_extended/metrics.py:117-123); NumPy stable argsort is
    the oracle.  Heavy ties (scores quantised to 1/64) and -0.0 / +0.0 included."""
    import ctypes
    from rodet_b200 import _abi
    rng = np.random.default_rng(n)
    scores = (np.round(rng.uniform(0, 1, n) * 64) / 64).astype(np.float32)
    scores[rng.uniform(size=n) < 0.05] = np.float32(-0.0)
    tp = (rng.uniform(size=n) < 0.4).astype(np.uint8)
    fp = (1 - tp).astype(np.uint8)
    order = np.argsort(-(scores + np.float32(0.0)), kind="stable")
    d = lambda a: torch.from_numpy(a).to(cuda_device)
    for k in sorted({n, max(1, n // 3)}):
        s_d, tp_d, fp_d = d(scores), d(tp), d(fp)
        o_tp, o_fp = torch.empty(k, dtype=torch.uint8, device=cuda_device), torch.empty(k, dtype=torch.uint8, device=cuda_device)
        o_s, o_i = torch.empty(k, dtype=torch.float32, device=cuda_device), torch.empty(k, dtype=torch.int32, device=cuda_device)
        nb = int(_abi.lib.rod_sort_scores_workspace_bytes(n))
        ws = torch.empty(nb, dtype=torch.uint8, device=cuda_device)
        _abi.check(_abi.lib.rod_sort_scores_desc(s_d.data_ptr(), n, k, tp_d.data_ptr(), fp_d.data_ptr(), o_tp.data_ptr(), o_fp.data_ptr(),
                                                 o_s.data_ptr(), o_i.data_ptr(), ws.data_ptr(), nb,
                                                 torch.cuda.current_stream(cuda_device).cuda_stream))
        assert np.array_equal(o_i.cpu().numpy(), order[:k].astype(np.int32))
        assert np.array_equal(o_tp.cpu().numpy(), tp[order[:k]]) and np.array_equal(o_fp.cpu().numpy(), fp[order[:k]])
        assert np.array_equal(o_s.cpu().numpy(), scores[order[:k]])
