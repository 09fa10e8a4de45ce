"""CPU-side checks (no GPU compute): the C-ABI library loads and exports every symbol the
header declares, the host mirror keeps the reference's names / signatures / error behaviour,
and the NumPy anchor functions reproduce the reference fixtures."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from conftest import LAYOUTS, ROOT, golden, golden_anchors

HEADER = os.path.join(ROOT, "include", "rodet_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rod_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import rodet_b200
    from rodet_b200 import _abi
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(_abi.lib, n), "librodet_b200.so does not export %s" % n
        assert n in _abi.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_abi.SIGNATURES) == names
    assert _abi.lib.rod_version() == 1
    assert isinstance(_abi.lib.rod_last_error(), bytes)


def test_struct_layout_matches_header():
    from rodet_b200 import _abi
    assert ctypes.sizeof(_abi.Layout) == 4 * (2 + 9)
    assert ctypes.sizeof(_abi.Layered) == 8 * 8 + 8 * 8


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No CPU fallback: importing the package without the .so raises ImportError."""
    import importlib.util
    import shutil
    pkg_src = os.path.join(ROOT, "road-object-detection-for-bdd100k_b200")
    dst = tmp_path / "pkgcopy"
    shutil.copytree(pkg_src, dst, ignore=shutil.ignore_patterns("*.so", "csrc", "__pycache__"))
    spec = importlib.util.spec_from_file_location("pkgcopy", dst / "__init__.py",
                                                  submodule_search_locations=[str(dst)])
    mod = importlib.util.module_from_spec(spec)
    import sys
    monkeypatch.setitem(sys.modules, "pkgcopy", mod)
    with pytest.raises(ImportError, match="no CPU fallback"):
        spec.loader.exec_module(mod)


def test_cpu_tensors_are_rejected():
    import torch
    from rodet_b200 import config
    from rodet_b200.utils import common_tools, net_tools
    import rodet_b200.utils.tf_extended as tfe
    with pytest.raises(ValueError, match="no CPU path"):
        common_tools.centerBboxes_2_cornerBboxes(torch.zeros(3, 4))
    with pytest.raises(ValueError, match="no CPU path"):
        net_tools.refine_groundtruth(golden_anchors("tiny"), torch.zeros(2, 4), torch.zeros(2, dtype=torch.int64),
                                     config.refine_method.JACCARD_BIGGER)
    with pytest.raises(ValueError, match="no CPU path"):
        tfe.bboxes_sort(torch.zeros(1, 8), torch.zeros(1, 8, 4), top_k=4)


def test_reference_error_behaviour():
    import torch
    from rodet_b200 import config
    from rodet_b200.utils import net_tools
    a = golden_anchors("tiny")
    with pytest.raises(ValueError, match="Not support now"):            # utils/net_tools.py:424
        net_tools.refine_groundtruth(a, torch.zeros(2, 4), torch.zeros(2), config.refine_method.JACCARD_TOPK)
    with pytest.raises(ValueError, match='Function parameter "method" wrong'):   # :426
        net_tools.refine_groundtruth(a, torch.zeros(2, 4), torch.zeros(2), "bogus")


def test_c_abi_argument_validation_without_gpu():
    """Entry points validate before touching the device: bad layouts / NULLs give ROD_E_INVALID."""
    from rodet_b200 import _abi
    lay = _abi.Layout()
    lay.n_layers = 0
    rc = _abi.lib.rod_arm_match_encode(lay, None, None, None, None, None, 1, None, 1, 1, 1, None, None, None, None, None, None)
    assert rc == _abi.E_INVALID and b"n_layers" in _abi.lib.rod_last_error()
    lay.n_layers, lay.n_total = 1, 8
    lay.offset[0], lay.offset[1] = 0, 8
    rc = _abi.lib.rod_arm_match_encode(lay, None, None, None, None, None, 1, None, 1, 1, 2, None, None, None, None, None, None)
    assert rc == _abi.E_UNSUPPORTED and _abi.lib.rod_last_error() == b"Not support now"
    rc = _abi.lib.rod_arm_match_encode(lay, None, None, None, None, None, 1, None, 1, 1, 1, None, None, None, None, None, None)
    assert rc == _abi.E_INVALID
    with pytest.raises(ValueError):
        _abi.check(rc)
    assert _abi.lib.rod_bboxes_sort(None, None, 1, 8, 4, None, None, None, None) == _abi.E_INVALID


def test_signatures_match_reference():
    """Same parameter names, order and defaults as the reference functions."""
    from rodet_b200.utils import common_tools, net_tools
    import rodet_b200.utils.tf_extended as tfe

    def params(fn, n=None):
        ps = list(inspect.signature(fn).parameters.values())
        return [(p.name, p.default) for p in (ps[:n] if n else ps)]
    E = inspect.Parameter.empty
    assert params(net_tools.init_anchor) == [("n_layers", E)]
    assert params(net_tools.anchors_one_layer) == [("img_shape", E), ("feat_shape", E), ("anchors_one_layer", E), ("dtype", np.float32)]
    assert params(net_tools.anchors_all_layer) == [("img_shape", E), ("feats_shape", E), ("anchors_all_layer", E)]
    assert params(net_tools.encode_locations_one_layer) == [("anchors_one_layer", E), ("center_bbox", E)]
    assert params(net_tools.decode_locations_one_layer) == [("anchors_one_layer", E), ("offset_bboxes", E)]
    assert params(net_tools.jaccard) == [("anchors", E), ("corner_bbox", E)]
    assert params(net_tools.refine_groundtruth, 5) == [("anchors_all_layer", E), ("center_bboxes", E), ("labels", E), ("method", E), ("scope", "refine_encode")]
    assert params(net_tools.det_groundtruth, 7) == [("refine_out", E), ("offset_gt", E), ("cbboxes", E), ("refine_labels", E), ("refine_pos_mask", E), ("anchors", E), ("scope", "det_encode")]
    sel = [("select_threshold", None), ("num_classes", 21), ("ignore_class", 0), ("scope", None)]
    assert params(net_tools.bboxes_select_one_layer) == [("predictions_layer", E), ("localizations_layer", E)] + sel
    assert params(net_tools.bboxes_select_all_layers) == [("predictions_net", E), ("localizations_net", E)] + sel
    assert params(net_tools.detected_bboxes, 7) == [("predictions", E), ("localisations", E), ("select_threshold", None), ("nms_threshold", 0.5), ("clipping_bbox", None), ("top_k", 800), ("keep_top_k", 200)]
    assert params(tfe.bboxes_sort) == [("scores", E), ("bboxes", E), ("top_k", 400), ("scope", None)]
    assert params(tfe.bboxes_sort_all_classes) == [("classes", E), ("scores", E), ("bboxes", E), ("top_k", 400), ("scope", None)]
    assert params(tfe.bboxes_nms) == [("scores", E), ("bboxes", E), ("nms_threshold", 0.5), ("keep_top_k", 200), ("scope", None)]
    assert params(tfe.bboxes_nms_batch) == [("scores", E), ("bboxes", E), ("nms_threshold", 0.5), ("keep_top_k", 200), ("scope", None)]
    assert params(tfe.bboxes_clip) == [("bbox_ref", E), ("bboxes", E), ("scope", None)]
    assert params(tfe.bboxes_resize) == [("bbox_ref", E), ("bboxes", E), ("name", None)]
    assert params(tfe.bboxes_jaccard) == [("bbox_ref", E), ("bboxes", E), ("name", None)]
    assert params(tfe.bboxes_intersection) == [("bbox_ref", E), ("bboxes", E), ("name", None)]
    assert params(tfe.pad_axis) == [("x", E), ("offset", E), ("size", E), ("axis", 0), ("name", None)]
    assert params(tfe.safe_divide) == [("numerator", E), ("denominator", E), ("name", None)]
    assert params(common_tools.centerBboxes_2_cornerBboxes) == [("center_bboxes", E)]
    assert params(common_tools.cornerBboxes_2_centerBboxes) == [("corner_bboxes", E)]
    # the "next" rows of SURVEY.md section 8 (f-1 .. f-4)
    import torch
    from rodet_b200.utils.tf_extended import tf_utils
    assert params(tfe.bboxes_matching, 7) == [("label", E), ("scores", E), ("bboxes", E), ("glabels", E), ("gbboxes", E), ("gdifficults", E), ("matching_threshold", 0.5)]
    assert params(tfe.bboxes_matching_batch, 7) == [("labels", E), ("scores", E), ("bboxes", E), ("glabels", E), ("gbboxes", E), ("gdifficults", E), ("matching_threshold", 0.5)]
    assert params(tfe.streaming_tp_fp_arrays, 8) == [("num_gbboxes", E), ("tp", E), ("fp", E), ("scores", E), ("remove_zero_scores", True), ("metrics_collections", None), ("updates_collections", None), ("name", None)]
    assert params(tfe.precision_recall) == [("num_gbboxes", E), ("num_detections", E), ("tp", E), ("fp", E), ("scores", E), ("dtype", torch.float64), ("scope", None)]
    assert params(tfe.average_precision_voc07) == [("precision", E), ("recall", E), ("name", None)]
    assert params(tfe.average_precision_voc12) == [("precision", E), ("recall", E), ("name", None)]
    assert params(tfe.precision_recall_values) == [("xvals", E), ("precision", E), ("recall", E), ("name", None)]
    assert params(tfe.cummax) == [("x", E), ("reverse", False), ("name", None)]
    assert params(tfe.bboxes_filter_overlap) == [("labels", E), ("bboxes", E), ("threshold", 0.5), ("assign_negative", False), ("scope", None)]
    assert params(tf_utils.reshape_list) == [("l", E), ("shape", None)]
    assert params(net_tools.smooth_l1) == [("x", E)]
    assert params(net_tools.refine_loss) == [("refine_out", E), ("refine_groundtruth", E), ("refine_pos_mask", E), ("dtype", torch.float32)]
    assert params(net_tools.det_clf_loss, 8) == [("refine_out", E), ("clf_out", E), ("det_out", E), ("det_groundtruth", E), ("det_pos_mask", E), ("det_labels", E), ("iou_all_layers", E), ("dtype", torch.float32)]


@pytest.mark.parametrize("layout", ["418", "512", "tiny"])
def test_host_anchor_functions_match_reference_fixture(layout):
    from rodet_b200 import config
    from rodet_b200.utils import net_tools
    img, feats = LAYOUTS[layout]
    config.img_size = img
    sizes = net_tools.init_anchor(6)
    assert np.array_equal(np.concatenate(list(sizes.values())), golden("anchors.npz")["%s_sizes_px" % layout])
    ours = net_tools.anchors_all_layer(img, {"layer_%d" % (i + 1): f for i, f in enumerate(feats)}, sizes)
    for o, r in zip(ours, golden_anchors(layout)):
        for a, b in zip(o, r):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    assert net_tools.n_anchor_each_layer("mobilenet_v2") == [6, 9, 9, 9, 9, 9]
    config.img_size = (418, 418)


def test_config_values_match_reference():
    from rodet_b200 import config
    assert config.refine_pos_jac_val_all_layers == [0.2, 0.3, 0.4, 0.4, 0.3, 0.3]
    assert config.det_pos_jac_val_all_layers == [0.5, 0.6, 0.7, 0.7, 0.6, 0.6]
    assert config.total_obj_n == 11 and config.img_size == (418, 418)
    assert config.normal_anchor_range == [0.05, 0.7] and config.special_anchor_range == [0.02, 0.03]
    assert [m.value for m in config.refine_method] == [0, 1, 2]
    assert config.feat_size_all_layers["mobilenet_v2"]["layer_1"] == (53, 53)


def test_synth_is_deterministic_and_in_domain():
    from rodet_b200 import synth
    a, la, ca = synth.gt_batch(10, 4)
    b, lb, cb = synth.gt_batch(10, 4)
    assert np.array_equal(a, b) and np.array_equal(la, lb) and np.array_equal(ca, cb)
    assert (ca >= 1).all() and (ca <= 100).all()
    for i in range(4):
        g = a[i, :ca[i]]
        assert ((g[:, 2] - g[:, 0]) >= 1e-3).all() and ((g[:, 3] - g[:, 1]) >= 1e-3).all()
        assert (g >= 0).all() and (g <= 1).all()
        assert (la[i, :ca[i]] >= 1).all() and (la[i, :ca[i]] <= 10).all()
    p = synth.class_probs(3, 500)
    assert np.allclose(p.sum(1), 1, atol=1e-5) and p.dtype == np.float32
    s = synth.stress_probs(3, 500)
    assert (s >= np.float32(0.3)).all()


def test_catch_net_layout_helpers():
    """nets/catch_net.py:306-308, 339-341, 344-363: head layout views and get_output's selection rule."""
    import torch
    from rodet_b200 import config
    from rodet_b200.nets import catch_net
    det_s, clf_s = catch_net.head_shapes(2)
    assert det_s[0] == (2, 53, 53, 6, 4) and clf_s[5] == (2, 2, 2, 9, 11)
    assert sum(s[1] * s[2] * s[3] for s in det_s) == 25800
    det_c = [torch.arange(2 * fh * fw * a * 4, dtype=torch.float32).reshape(2, fh, fw, a * 4) for (_, fh, fw, a, _) in det_s]
    clf_c = [torch.zeros(2, fh, fw, a * 11) for (_, fh, fw, a, _) in det_s]
    net = catch_net.factory(det_c, det_c, clf_c, config_dict={'train_range': config.train_range.ALL})
    ro, do, co = net.get_output()
    assert [tuple(t.shape) for t in ro] == det_s and [tuple(t.shape) for t in co] == clf_s
    assert ro[0].data_ptr() == det_c[0].data_ptr()                       # views, not copies
    assert ro[1][1, 3, 4, 2, 1] == det_c[1][1, 3, 4, 2 * 4 + 1]          # channel = a * 4 + k
    only = catch_net.factory(det_c, config_dict={'train_range': config.train_range.REFINE}).get_output()
    assert len(only) == 6 and tuple(only[2].shape) == det_s[2]
    with pytest.raises(ValueError):
        catch_net.det_out(det_c[:3])
    with pytest.raises(ValueError):
        catch_net.clf_out(det_c)
    bad = catch_net.factory(det_c, det_c, clf_c)
    bad.train_range = None
    with pytest.raises(ValueError, match='Error'):
        bad.get_output()


def test_reshape_list_round_trip():
    """utils/tf_extended/tf_utils.py:29-55 as used by train.py:114-124."""
    from rodet_b200.utils.tf_extended import tf_utils
    img, a, b = "img", ["a%d" % i for i in range(6)], ("b%d" % i for i in range(6))
    b = list(b)
    flat = tf_utils.reshape_list([img, a, tuple(b)])
    assert flat == [img] + a + b
    assert tf_utils.reshape_list(flat, [1, 6, 6]) == [img, a, b]
    assert tf_utils.reshape_list([], None) == [] and tf_utils.reshape_list([1, 2], [1, 1]) == [1, 2]


def test_loss_layout_from_lists_and_metric_constants():
    """Host-side pieces of f-2 / f-3 that need no GPU: the layer layout derived from per-layer head tensors, the VOC07
    recall levels (np.arange(0., 1.1, 0.1) bit for bit), smooth_l1's formula."""
    import torch
    from rodet_b200.utils import losses
    from rodet_b200.utils.tf_extended import metrics
    ts = [torch.zeros(2, 4, 4, 6, 11), torch.zeros(2, 2, 2, 9, 11), torch.zeros(2, 1, 1, 9, 11)]
    sh = losses._Shapes(ts, 11)
    assert sh.n == 4 * 4 * 6 + 2 * 2 * 9 + 9 and sh.offsets == [0, 96, 132, 141] and sh.n_layers == 3
    assert sh.layout.n_total == 141 and [sh.layout.offset[i] for i in range(4)] == [0, 96, 132, 141]
    flat = torch.arange(2 * 141 * 11, dtype=torch.float32).reshape(2, 141, 11)
    parts = sh.split(flat, 11)
    assert [tuple(p.shape) for p in parts] == [tuple(t.shape) for t in ts] and parts[1][1, 1, 0, 3, 5] == flat[1, 96 + (1 * 2 + 0) * 9 + 3, 5]
    with pytest.raises(ValueError):
        losses._Shapes([torch.zeros(2, 4, 4, 6, 4)], 11)
    assert np.array_equal(metrics.VOC07_RECALL_LEVELS, np.arange(0., 1.1, 0.1)) and metrics.VOC07_RECALL_LEVELS.size == 11
    x = torch.tensor([-2.0, -0.5, 0.0, 0.25, 1.0, 3.0])
    assert torch.allclose(losses.smooth_l1(x), torch.tensor([1.5, 0.125, 0.0, 0.03125, 0.5, 2.5]))


def test_layer_lists_are_real_lists():
    """The per-layer lists returned by refine_groundtruth / det_groundtruth are consumed by torch.cat / torch.stack /
    list concatenation, which read the list storage directly (the reference returns plain lists)."""
    import torch
    from rodet_b200 import _abi
    from rodet_b200.anchor_table import AnchorTable
    shapes = [(3, 2, 6), (2, 2, 9), (1, 1, 9)]
    n = sum(a * b * c for a, b, c in shapes)
    table = AnchorTable(shapes, torch.zeros(n, 4), torch.zeros(n, 4))
    flat4 = torch.arange(2 * n * 4, dtype=torch.float32).view(2, n, 4)
    flat1 = torch.arange(2 * n, dtype=torch.int32).view(2, n)
    l4 = _abi.LayerList(flat4, table, True, False)
    l1 = _abi.LayerList(flat1, table, True, True)
    assert list.__len__(l4) == 3 and len(l4 + [1]) == 4 and isinstance(l4, list)
    assert [tuple(t.shape) for t in l4] == [(2, 3, 2, 6, 4), (2, 2, 2, 9, 4), (2, 1, 1, 9, 4)]
    assert [tuple(t.shape) for t in l1] == [(2, 3, 2, 6, 1), (2, 2, 2, 9, 1), (2, 1, 1, 9, 1)]
    cat = torch.cat([t.reshape(2, -1, 4) for t in l4], 1)
    assert torch.equal(cat, flat4)
    assert torch.equal(torch.cat([t.reshape(2, -1) for t in l1], 1), flat1)
    assert torch.stack(list(l4[1:2])).shape == (1, 2, 2, 2, 9, 4)
    assert l4[0].data_ptr() == flat4.data_ptr()                  # zero-copy views
    one = _abi.LayerList(flat4[:1], table, False, False)          # per-image (reference) form: no batch axis
    assert [tuple(t.shape) for t in one] == [(3, 2, 6, 4), (2, 2, 9, 4), (1, 1, 9, 4)]
    idx = _abi.LayerList(flat1, table, True, False)
    assert [tuple(t.shape) for t in idx] == [(2, 3, 2, 6), (2, 2, 2, 9), (2, 1, 1, 9)]
    assert torch.equal(idx[1], flat1[:, 36:72].view(2, 2, 2, 9))


def test_anchor_cache_key_follows_content():
    from rodet_b200 import anchor_table
    import numpy as np
    a = [[np.zeros((2, 2, 1), np.float32), np.ones((2, 2, 1), np.float32), np.ones(3, np.float32), np.ones(3, np.float32)]]
    b = [[v.copy() for v in a[0]]]
    k1, k2 = anchor_table._content_key(a, "cuda:0", "all"), anchor_table._content_key(b, "cuda:0", "all")
    assert k1 == k2                                               # equal anchors rebuilt by the caller hit the cache
    b[0][2][1] = 2.0                                              # in-place mutation changes the key
    assert anchor_table._content_key(b, "cuda:0", "all") != k1
    assert anchor_table._content_key(a, "cuda:1", "all") != k1 and anchor_table._content_key(a, "cuda:0", "one") != k1


def test_crop_without_overlap_threshold_is_rejected():
    """process_raw_gt_train: the reference always filters a crop's boxes at 0.3 (utils/augmentation/process.py:134-138);
    a crop with crop_overlap=None has no reference meaning and must not silently filter at 0."""
    import torch
    from rodet_b200.utils.data_pileline_tools import process_raw_gt_train
    labels, boxes = torch.zeros((1, 2), dtype=torch.int64), torch.zeros((1, 2, 4))
    with pytest.raises(ValueError, match="crop_overlap=None"):
        process_raw_gt_train(labels, boxes, None, torch.tensor([[0.0, 0.0, 1.0, 1.0]]), None, crop_overlap=None)
