"""GPU tests at the C-ABI level: raw-pointer entry points called directly through ctypes, DLPack
validation errors, unsupported-parameter error codes."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_anchors
from helpers import bit_equal
from oracle import restated as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(cuda_device):
    import rodet_b200
    from rodet_b200 import _abi, synth
    from rodet_b200.utils import net_tools

    class NS:
        pass
    ns = NS()
    ns.abi, ns.synth, ns.nt, ns.dev = _abi, synth, net_tools, cuda_device
    ns.anchors = golden_anchors("418")
    ns.table = rodet_b200.AnchorTable.from_anchors(ns.anchors, cuda_device)
    ns.otable = R.AnchorTable(ns.anchors)
    return ns


def test_raw_pointer_arm_and_odm(ctx):
    """rod_arm_match_encode / rod_odm_target with plain device pointers and host threshold arrays."""
    abi, t, B = ctx.abi, ctx.table, 3
    corner, labels, counts = ctx.synth.gt_batch(60, B)
    center = R.corner_to_center(corner).astype(np.float32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.dev)
    cb_d, lab_d, cnt_d = d(center), d(labels.astype(np.int32)), d(counts)       # int32 labels this time
    N = t.n
    gt = torch.empty((B, N, 4), device=ctx.dev); cb = torch.empty((B, N, 4), device=ctx.dev)
    lb = torch.empty((B, N), dtype=torch.int32, device=ctx.dev); ps = torch.empty_like(lb); ix = torch.empty_like(lb)
    thr = abi.float_array(R.REFINE_POS_JAC)
    s = abi.stream_ptr(ctx.dev)
    rc = abi.lib.rod_arm_match_encode(t.layout, t.corner.data_ptr(), t.center.data_ptr(), thr, cb_d.data_ptr(),
                                      lab_d.data_ptr(), 0, cnt_d.data_ptr(), B, center.shape[1], 1, gt.data_ptr(),
                                      cb.data_ptr(), lb.data_ptr(), ps.data_ptr(), ix.data_ptr(), s)
    assert rc == 0, abi.lib.rod_last_error()
    for b in range(B):
        o = R.arm_match_encode(ctx.otable, center[b, :counts[b]], labels[b, :counts[b]])
        assert np.array_equal(ps[b].cpu().numpy(), o[3]) and np.array_equal(ix[b].cpu().numpy(), o[4])
        assert bit_equal(gt[b].cpu().numpy(), o[0]) and np.array_equal(lb[b].cpu().numpy(), o[2])
    # ODM through rod_layered_t built by hand from the flat ARM outputs
    ro = d(np.stack([ctx.synth.head_offsets(60 + b, N) for b in range(B)]))

    def lay(flat, inner):
        L = abi.Layered()
        for l in range(t.n_layers):
            L.base[l] = flat.data_ptr() + t.offsets[l] * inner * flat.element_size()
            L.batch_stride[l] = flat.stride(0)
        return L
    det = torch.empty((B, N, 4), device=ctx.dev); mk = torch.empty_like(lb); dl = torch.empty_like(lb)
    iou = torch.empty((B, N), device=ctx.dev)
    rc = abi.lib.rod_odm_target(t.layout, t.center.data_ptr(), abi.float_array(R.DET_POS_JAC), lay(ro, 4), lay(gt, 4),
                                lay(cb, 4), lay(lb, 1), lay(ps, 1), B, det.data_ptr(), mk.data_ptr(), dl.data_ptr(),
                                iou.data_ptr(), s)
    assert rc == 0, abi.lib.rod_last_error()
    o = R.odm_target(ctx.otable, ro.cpu().numpy(), gt.cpu().numpy(), cb.cpu().numpy(), lb.cpu().numpy(), ps.cpu().numpy())
    assert np.array_equal(mk.cpu().numpy(), o[1]) and bit_equal(iou.cpu().numpy(), o[3]) and bit_equal(det.cpu().numpy(), o[0])


def test_dlpack_validation_errors(ctx):
    nt, dev = ctx.nt, ctx.dev
    from rodet_b200 import config
    JB = config.refine_method.JACCARD_BIGGER
    good_b = torch.rand(2, 5, 4, device=dev) * 0.2 + 0.3
    good_l = torch.ones(2, 5, dtype=torch.int64, device=dev)
    with pytest.raises(ValueError, match="float32"):
        nt.refine_groundtruth(ctx.anchors, good_b.double(), good_l, JB)
    with pytest.raises(ValueError, match="labels"):
        nt.refine_groundtruth(ctx.anchors, good_b, good_l.float(), JB)
    with pytest.raises(ValueError, match="do not agree"):
        nt.refine_groundtruth(ctx.anchors, good_b, good_l[:, :3], JB)
    # a per-layer list with the wrong number of anchors is rejected by the C-side DLPack check
    t = ctx.table
    ro = [torch.zeros((2,) + s + (4,), device=dev) for s in t.shapes]
    bad = list(ro)
    bad[2] = torch.zeros((2, 3, 3, 9, 4), device=dev)
    arm = nt.refine_groundtruth(ctx.anchors, good_b, good_l, JB)
    with pytest.raises(ValueError, match="elements per image"):
        nt.det_groundtruth(bad, arm[0], arm[1], arm[2], arm[3], ctx.anchors)
    with pytest.raises(ValueError, match="dtype"):
        nt.det_groundtruth([x.double() for x in ro], arm[0], arm[1], arm[2], arm[3], ctx.anchors)
    out = nt.det_groundtruth(ro, arm[0], arm[1], arm[2], arm[3], ctx.anchors)      # and the good call works
    assert out[3][0].shape == (2,) + t.shapes[0]


def test_unsupported_and_invalid_parameters(ctx):
    nt, dev, t = ctx.nt, ctx.dev, ctx.table
    preds = [torch.rand((1,) + s + (11,), device=dev) for s in t.shapes]
    locs = [torch.rand((1,) + s + (4,), device=dev) for s in t.shapes]
    with pytest.raises(ValueError, match="not supported"):
        nt.detected_bboxes(preds, locs, select_threshold=0.3, top_k=2000)
    with pytest.raises(ValueError, match="top_k"):
        nt.detected_bboxes(preds, locs, select_threshold=0.3, top_k=10 ** 6)
    import rodet_b200.utils.tf_extended as tfe
    with pytest.raises(ValueError, match="top_k"):
        tfe.bboxes_sort(torch.rand(2, 10, device=dev), torch.rand(2, 10, 4, device=dev), top_k=11)
    with pytest.raises(ValueError, match="not supported"):
        tfe.bboxes_nms_batch(torch.rand(1, 2000, device=dev), torch.rand(1, 2000, 4, device=dev))
    # workspace too small / misaligned at the raw ABI
    abi = ctx.abi
    need = abi.lib.rod_detect_workspace_bytes(t.layout, 1, 11, 400)
    assert need > 0
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    sc = torch.empty((11, 1, 200), device=dev); bx = torch.empty((11, 1, 200, 4), device=dev)
    pl, k1 = abi.layered(preds, 11)
    ll, k2 = abi.layered(locs, 4)
    rc = abi.lib.rod_detect(t.layout, None, pl, ll, None, None, 1, 11, 0, 0.3, 0.45, 400, 200, None, sc.data_ptr(),
                            bx.data_ptr(), None, ws.data_ptr(), 1024, abi.stream_ptr(dev))
    assert rc == abi.E_INVALID and b"workspace" in abi.lib.rod_last_error()
