"""Guard-band tests of the two hot-path entry points at the raw C-ABI level (compute-sanitizer is not
available on the GPU pool, so out-of-bounds WRITES are looked for directly): every output buffer and the
detect workspace are carved out of a larger allocation filled with a byte pattern, at exactly the size the
header asks for; after the call the bytes in front of and behind each buffer must be untouched and the
results must equal those of the public Python API on ordinary allocations.  Geometry: the 512 x 512 layout of
the bench (TMA scan, sampled cuts, spill lists), small batches, score workloads that fill the list slices."""
import numpy as np
import pytest
import torch

from conftest import LAYOUTS, golden_anchors

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5


class Guarded:
    """`nbytes` of device memory with GUARD pattern bytes on both sides (the payload is pattern-filled as well)."""

    def __init__(self, nbytes, dev, align=256):
        self.n = int(nbytes)
        self.raw = torch.full((self.n + 2 * GUARD + align,), PATTERN, dtype=torch.uint8, device=dev)
        off = GUARD + (-(self.raw.data_ptr() + GUARD)) % align
        self.off = off
        self.view = self.raw[off:off + self.n]
        assert self.view.data_ptr() % align == 0

    def ptr(self):
        return self.view.data_ptr()

    def as_tensor(self, dtype, shape):
        return self.view.view(dtype).view(shape)

    def intact(self):
        front, back = self.raw[:self.off], self.raw[self.off + self.n:]
        return bool((front == PATTERN).all()) and bool((back == PATTERN).all())


@pytest.fixture(scope="module")
def ctx(cuda_device):
    import rodet_b200
    from rodet_b200 import _abi, config, synth
    from rodet_b200.utils import net_tools

    class NS:
        pass
    ns = NS()
    ns.abi, ns.synth, ns.nt, ns.dev, ns.config = _abi, synth, net_tools, cuda_device, config
    ns.anchors = golden_anchors("512")
    ns.table = rodet_b200.AnchorTable.from_anchors(ns.anchors, cuda_device)
    ns.shapes = [(fh, fw, 6 if i == 0 else 9) for i, (fh, fw) in enumerate(LAYOUTS["512"][1])]      # (fh, fw, anchors per cell)
    assert sum(a * b * c for a, b, c in ns.shapes) == ns.table.n
    return ns


def _layered(ctx, flat, inner):
    t, L = ctx.table, ctx.abi.Layered()
    for l in range(t.n_layers):
        L.base[l] = flat.data_ptr() + t.offsets[l] * inner * flat.element_size()
        L.batch_stride[l] = flat.stride(0)
    return L


def _split(ctx, flat):
    t = ctx.table
    return [flat[:, t.offsets[l]:t.offsets[l + 1]].contiguous() for l in range(t.n_layers)]


def test_target_fused_guard_bands(ctx):
    """rod_target_fused (train.py:109-113 -> :147-149): nine outputs + the 8-byte scheduler workspace, each with
    guard bands; GT counts 0, 1 and Gmax are in the batch."""
    abi, t, dev, B = ctx.abi, ctx.table, ctx.dev, 4
    N = t.n
    corner, labels, counts = ctx.synth.gt_batch(7_000, B)
    counts = counts.copy()
    counts[0], counts[1], counts[2] = 0, 1, corner.shape[1]
    center = np.concatenate([(corner[..., :2] + corner[..., 2:]) * 0.5, corner[..., 2:] - corner[..., :2]], -1).astype(np.float32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cb_d, lab_d, cnt_d = d(center), d(labels.astype(np.int64)), d(counts.astype(np.int32))
    ro = d(np.stack([ctx.synth.head_offsets(7_000 + b, N) for b in range(B)]))
    sizes = {"gt": 16, "cb": 16, "lab": 4, "pos": 4, "idx": 4, "det_gt": 16, "mask": 4, "dlab": 4, "iou": 4}
    g = {k: Guarded(B * N * v, dev) for k, v in sizes.items()}
    ws = Guarded(int(abi.lib.rod_target_fused_workspace_bytes()), dev)
    ws.view.zero_()
    ta = abi.float_array(ctx.config.refine_pos_jac_val_all_layers)
    to = abi.float_array(ctx.config.det_pos_jac_val_all_layers)
    for rep in range(2):                                   # twice: the scheduler counters must be left at zero
        rc = abi.lib.rod_target_fused(t.layout, t.corner.data_ptr(), t.center.data_ptr(), ta, to, cb_d.data_ptr(), lab_d.data_ptr(), 1,
                                      cnt_d.data_ptr(), B, center.shape[1], _layered(ctx, ro, 4), g["gt"].ptr(), g["cb"].ptr(),
                                      g["lab"].ptr(), g["pos"].ptr(), g["idx"].ptr(), g["det_gt"].ptr(), g["mask"].ptr(),
                                      g["dlab"].ptr(), g["iou"].ptr(), ws.ptr(), abi.stream_ptr(dev))
        assert rc == 0, abi.lib.rod_last_error()
        torch.cuda.synchronize()
        assert int(ws.view.to(torch.int32).sum()) == 0, "scheduler workspace not reset"
        for k, b in g.items():
            assert b.intact(), "write outside the %s buffer" % k
        assert ws.intact(), "write outside the scheduler workspace"
    # same bits as the two-call public API on ordinary allocations
    JB = ctx.config.refine_method.JACCARD_BIGGER
    arm = ctx.nt.refine_groundtruth(t, cb_d, lab_d, JB, gt_counts=cnt_d, return_match_index=True)
    det = ctx.nt.det_groundtruth(_split(ctx, ro), arm[0], arm[1], arm[2], arm[3], t)
    f32, i32 = torch.float32, torch.int32
    want = {"gt": arm[0].flat, "cb": arm[1].flat, "lab": arm[2].flat, "pos": arm[3].flat, "idx": arm[4].flat,
            "det_gt": det[0].flat, "mask": det[1].flat, "dlab": det[2].flat, "iou": det[3].flat}
    for k, w in want.items():
        got = g[k].as_tensor(f32 if w.dtype == f32 else i32, w.shape)
        assert torch.equal(got.view(i32), w.contiguous().view(i32)), k


@pytest.mark.parametrize("kind,B", [("normal", 3), ("stress", 2), ("quadrant", 3), ("bumps", 2)])
def test_detect_guard_bands(ctx, kind, B):
    """rod_detect (evaluate.py:139-151) with outputs and a workspace of exactly rod_detect_workspace_bytes() between guard bands, on
    score workloads that fill the per-CTA list slices and the spill lists; results equal decode_detected_bboxes()."""
    abi, t, dev = ctx.abi, ctx.table, ctx.dev
    N, C, top_k, keep = t.n, 11, 400, 200
    if kind == "stress":
        probs = np.stack([ctx.synth.stress_probs(900 + b, N) for b in range(B)])
    elif kind == "normal":
        probs = np.stack([ctx.synth.class_probs(900 + b, N) for b in range(B)])
    else:
        probs = np.stack([ctx.synth.clustered_probs(900 + b, ctx.shapes, kind) for b in range(B)])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).to(dev)
    p = d(probs)
    ro = d(np.stack([ctx.synth.head_offsets(900 + b, N, 0) for b in range(B)]))
    do = d(np.stack([ctx.synth.head_offsets(900 + b, N, 1) for b in range(B)]))
    need = int(abi.lib.rod_detect_workspace_bytes(t.layout, B, C, top_k))
    sc, bx, cn, ws = Guarded(C * B * keep * 4, dev), Guarded(C * B * keep * 16, dev), Guarded(C * B * 4, dev), Guarded(need, dev)
    off = int(abi.lib.rod_detect_flags_offset(t.layout, B, C, top_k))
    ws.view[off:off + int(abi.lib.rod_detect_workspace_clean_bytes(t.layout, B, C, top_k))].zero_()
    for rep in range(2):                                   # second call: the self-cleaning workspace head
        rc = abi.lib.rod_detect(t.layout, t.center.data_ptr(), _layered(ctx, p, C), None, _layered(ctx, ro, 4), _layered(ctx, do, 4),
                                B, C, 0, 0.3, 0.45, top_k, keep, None, sc.ptr(), bx.ptr(), cn.ptr(), ws.ptr(), need, abi.stream_ptr(dev))
        assert rc == 0, abi.lib.rod_last_error()
        torch.cuda.synchronize()
        for name, b in (("scores", sc), ("boxes", bx), ("counts", cn), ("workspace", ws)):
            assert b.intact(), "write outside the %s buffer (%s)" % (name, kind)
    rs, rb, rc_ = ctx.nt.decode_detected_bboxes(t, _split(ctx, ro), _split(ctx, do), _split(ctx, p), select_threshold=0.3,
                                                nms_threshold=0.45, top_k=top_k, keep_top_k=keep, return_counts=True)
    i32 = torch.int32
    got_s = sc.as_tensor(torch.float32, (C, B, keep))
    got_b = bx.as_tensor(torch.float32, (C, B, keep, 4))
    got_c = cn.as_tensor(i32, (C, B))
    for c in range(1, C):
        assert torch.equal(got_s[c].view(i32), rs[c].view(i32)) and torch.equal(got_b[c].view(i32), rb[c].view(i32)), (kind, c)
    assert torch.equal(got_c[1:], rc_[1:])
    assert int(got_c[1:].sum()) > 0


def test_detect_generic_depth_guard_bands(ctx, monkeypatch):
    """The plain-load kernels (prediction depths other than 11: hist / thresh / collect) with the same guard bands."""
    abi, t, dev, B, C, top_k, keep = ctx.abi, ctx.table, ctx.dev, 2, 5, 300, 100
    N = t.n
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).to(dev)
    p = d(np.stack([ctx.synth.class_probs(40 + b, N, C) for b in range(B)]))
    loc = d(np.stack([ctx.synth.head_offsets(40 + b, N, 0) for b in range(B)]))
    need = int(abi.lib.rod_detect_workspace_bytes(t.layout, B, C, top_k))
    sc, bx, cn, ws = Guarded(C * B * keep * 4, dev), Guarded(C * B * keep * 16, dev), Guarded(C * B * 4, dev), Guarded(need, dev)
    rc = abi.lib.rod_detect(t.layout, None, _layered(ctx, p, C), _layered(ctx, loc, 4), None, None, B, C, 0, 0.05, 0.45, top_k, keep,
                            None, sc.ptr(), bx.ptr(), cn.ptr(), ws.ptr(), need, abi.stream_ptr(dev))
    assert rc == 0, abi.lib.rod_last_error()
    torch.cuda.synchronize()
    for name, b in (("scores", sc), ("boxes", bx), ("counts", cn), ("workspace", ws)):
        assert b.intact(), "write outside the %s buffer" % name
    monkeypatch.setattr(ctx.config, "total_obj_n", C)
    rs, rb = ctx.nt.detected_bboxes(_split(ctx, p), _split(ctx, loc), select_threshold=0.05, nms_threshold=0.45, top_k=top_k,
                                    keep_top_k=keep)
    i32 = torch.int32
    got_s, got_b = sc.as_tensor(torch.float32, (C, B, keep)), bx.as_tensor(torch.float32, (C, B, keep, 4))
    for c in range(1, C):
        assert torch.equal(got_s[c].view(i32), rs[c].view(i32)) and torch.equal(got_b[c].view(i32), rb[c].view(i32)), c
