"""f-3: refine_loss / det_clf_loss (utils/net_tools.py:478-623) on the GPU: values against the fixture written by
the unmodified reference and against oracle/restated on targets produced by our own ARM / ODM kernels (tolerance
1e-5, the north star's float tolerance); gradients against a plain PyTorch float64 implementation of the same
formula with the targets held constant."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_anchors
from oracle import restated as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


def close(a, b, tol=TOL):
    a = float(a.detach()) if torch.is_tensor(a) else float(a)
    return abs(a - float(b)) <= tol * max(abs(float(b)), 1e-12)


@pytest.fixture(scope="module")
def env(cuda_device):
    from rodet_b200 import config, synth
    from rodet_b200.utils import net_tools

    class NS:
        pass
    ns = NS()
    ns.nt, ns.synth, ns.config, ns.dev = net_tools, synth, config, cuda_device
    return ns


def test_losses_vs_reference_fixture(env):
    z = golden("losses.npz")
    L = int(z["n_layers"])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    get = lambda k: [d(z["%s_%d" % (k, l)]) for l in range(L)]
    rl = env.nt.refine_loss(get("ro"), get("refine_gt"), get("refine_pos"))
    dl, cl, det = env.nt.det_clf_loss(get("ro"), get("clf"), get("do"), get("det_gt"), get("det_mask"), get("det_lab"), get("iou"),
                                      return_details=True)
    assert rl.dim() == 0 and rl.dtype == torch.float32
    assert close(rl, z["refine_loss"]) and close(dl, z["det_loss"]) and close(cl, z["clf_loss"])
    o = R.clf_loss([z["clf_%d" % l] for l in range(L)], [z["det_mask_%d" % l] for l in range(L)],
                   [z["det_lab_%d" % l] for l in range(L)], [z["iou_%d" % l] for l in range(L)])
    assert int(det["n_pos"]) == o["n_pos"] and int(det["n_neg"]) == o["n_neg"]
    assert close(det["pos_loss"], o["pos_loss"]) and close(det["neg_loss"], o["neg_loss"])
    assert close(det["max_hard_pred"], o["max_hard_pred"], 1e-5)


def _pipeline(env, B, first):
    """ARM -> ODM targets from our kernels on BASELINE-shaped GT at 418x418, random heads."""
    anchors = golden_anchors("418")
    table = R.AnchorTable(anchors)
    corner, labels, counts = env.synth.gt_batch(first, B)
    center = R.corner_to_center(corner).astype(np.float32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    JB = env.config.refine_method.JACCARD_BIGGER
    gt, cb, lab, pos = env.nt.refine_groundtruth(anchors, d(center), d(labels), JB, gt_counts=d(counts))
    rng = np.random.default_rng(first)
    ro_flat = np.stack([env.synth.head_offsets(first + b, table.n) for b in range(B)])
    gflat = np.concatenate([t.cpu().numpy().reshape(B, -1, 4) for t in gt], axis=1)
    pflat = np.concatenate([t.cpu().numpy().reshape(B, -1) for t in pos], axis=1).astype(bool)
    ro_flat[pflat] = gflat[pflat] + (rng.standard_normal(gflat[pflat].shape) * 0.05).astype(np.float32)
    split = lambda flat, tail: [d(np.ascontiguousarray(flat[:, table.offsets[l]:table.offsets[l + 1]]).reshape((B,) + table.shapes[l] + tail))
                                for l in range(6)]
    ro = split(ro_flat, (4,))
    do = split(np.stack([env.synth.head_offsets(first + b, table.n, 1) for b in range(B)]), (4,))
    clf = split(np.stack([env.synth.class_logits(first + b, table.n) for b in range(B)]), (11,))
    det_gt, mask, dlab, iou = env.nt.det_groundtruth(ro, gt, cb, lab, pos, anchors)
    return dict(ro=ro, do=do, clf=clf, gt=gt, pos=pos, det_gt=det_gt, mask=mask, dlab=dlab, iou=iou, B=B)


def _np(ts):
    return [t.cpu().numpy() for t in ts]


@pytest.mark.parametrize("B,first", [(4, 7000), (8, 7100)])
def test_losses_on_pipeline_targets(env, B, first):
    p = _pipeline(env, B, first)
    rl = env.nt.refine_loss(p["ro"], p["gt"], p["pos"])                      # LayerList inputs straight from the ARM kernel
    dl, cl, det = env.nt.det_clf_loss(p["ro"], p["clf"], p["do"], p["det_gt"], p["mask"], p["dlab"], p["iou"], return_details=True)
    assert close(rl, R.smooth_l1_loss(_np(p["gt"]), _np(p["ro"]), _np(p["pos"])))
    assert close(dl, R.smooth_l1_loss(_np(p["det_gt"]), _np(p["do"]), _np(p["mask"])))
    o = R.clf_loss(_np(p["clf"]), _np(p["mask"]), _np(p["dlab"]), _np(p["iou"]))
    assert o["n_pos"] > 20 and int(det["n_pos"]) == o["n_pos"] and int(det["n_neg"]) == o["n_neg"] == 3 * o["n_pos"] + B
    assert close(det["pos_loss"], o["pos_loss"]) and close(det["neg_loss"], o["neg_loss"], 1e-4) and close(cl, o["clf_loss"], 1e-4)
    # the same through plain per-layer tensors (a caller that does not use our LayerLists)
    plain = lambda ts: [t.clone() for t in ts]
    dl2, cl2 = env.nt.det_clf_loss(p["ro"], p["clf"], p["do"], plain(p["det_gt"]), plain(p["mask"]), plain(p["dlab"]), plain(p["iou"]))
    assert float(dl2) == float(dl) and float(cl2) == float(cl)


def test_loss_gradients_vs_torch(env):
    """backward() of the three losses against PyTorch autograd on the same formula in float64 (targets constant)."""
    p = _pipeline(env, 4, 7200)
    B = p["B"]
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    ro, do, clf = leaf(p["ro"]), leaf(p["do"]), leaf(p["clf"])
    rl = env.nt.refine_loss(ro, p["gt"], p["pos"])
    dl, cl = env.nt.det_clf_loss(ro, clf, do, p["det_gt"], p["mask"], p["dlab"], p["iou"])
    (rl * 2.0 + dl * 0.5 + cl * 3.0).backward()

    def sl1(y, x, m):
        z = (y.double() - x) * m.double()
        a = z.abs()
        return (0.5 * ((a - 1) * torch.clamp(a, max=1) + a)).sum() / B
    ro64, do64, clf64 = [t.detach().double().requires_grad_(True) for t in ro], [t.detach().double().requires_grad_(True) for t in do], \
        [t.detach().double().requires_grad_(True) for t in clf]
    ref_rl = sum(sl1(y, x, m) for y, x, m in zip(p["gt"], ro64, p["pos"]))
    ref_dl = sum(sl1(y, x, m) for y, x, m in zip(p["det_gt"], do64, p["mask"]))
    o = R.clf_loss(_np(p["clf"]), _np(p["mask"]), _np(p["dlab"]), _np(p["iou"]))
    w = torch.from_numpy(o["weights"]).to(env.dev)
    tgt = torch.from_numpy(o["targets"]).to(env.dev)
    logits = torch.cat([t.reshape(-1, 11) for t in clf64], 0)
    ce = torch.nn.functional.cross_entropy(logits, tgt, reduction="none")
    ref_cl = (ce * w).sum()
    (ref_rl * 2.0 + ref_dl * 0.5 + ref_cl * 3.0).backward()
    for ours, ref in ((ro, ro64), (do, do64), (clf, clf64)):
        for a, b in zip(ours, ref):
            assert a.grad is not None and a.grad.shape == a.shape
            err = (a.grad.double() - b.grad).abs().max().item()
            scale = b.grad.abs().max().item()
            assert err <= 2e-6 * max(scale, 1e-6) + 1e-9, (err, scale)
    assert close(cl, ref_cl.item(), 1e-4)


def test_reference_gradients_reach_refine_out(env):
    """reference_gradients=True: det_loss also back-propagates into refine_out through
    det_gt = (offset_gt - refine_out) * mask (utils/net_tools.py:471; the reference has no stop_gradient), checked
    against PyTorch autograd on the reference's formula in float64 with det_gt rebuilt from refine_out."""
    p = _pipeline(env, 4, 7300)
    B = p["B"]
    leaf = lambda ts: [t.clone().requires_grad_(True) for t in ts]
    ro, do = leaf(p["ro"]), leaf(p["do"])
    dl, _ = env.nt.det_clf_loss(ro, p["clf"], do, p["det_gt"], p["mask"], p["dlab"], p["iou"], reference_gradients=True)
    dl.backward()
    ro64 = [t.detach().double().requires_grad_(True) for t in ro]
    do64 = [t.detach().double().requires_grad_(True) for t in do]
    ref = 0.0
    for og, r, d, m in zip(p["gt"], ro64, do64, p["mask"]):
        mf = m.double()
        det_gt = (og.double() - r) * mf                       # :471, differentiable w.r.t. refine_out
        z = (det_gt - d) * mf
        a = z.abs()
        ref = ref + (0.5 * ((a - 1) * torch.clamp(a, max=1) + a)).sum() / B
    ref.backward()
    n_nonzero = 0
    for ours, r64 in ((ro, ro64), (do, do64)):
        for a, b in zip(ours, r64):
            assert a.grad is not None
            err = (a.grad.double() - b.grad).abs().max().item()
            assert err <= 2e-6 * max(b.grad.abs().max().item(), 1e-6) + 1e-9
            n_nonzero += int((a.grad != 0).sum())
    assert n_nonzero > 0
    for r, d in zip(ro, do):
        assert torch.equal(r.grad, d.grad)
    # default mode: refine_out receives nothing from det_loss
    ro2, do2 = leaf(p["ro"]), leaf(p["do"])
    dl2, _ = env.nt.det_clf_loss(ro2, p["clf"], do2, p["det_gt"], p["mask"], p["dlab"], p["iou"])
    dl2.backward()
    assert float(dl2.detach()) == float(dl.detach()) and all(t.grad is None for t in ro2)
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(do, do2))


def test_invalid_label_gives_nan_not_garbage(env):
    """A positive anchor with a label outside [0, C): NaN loss / gradient (TF's GPU kernel), never an OOB read."""
    p = _pipeline(env, 2, 7400)
    lab = [t.clone() for t in p["dlab"]]
    mask = [t.clone() for t in p["mask"]]
    lab[1].view(-1)[5] = 11
    mask[1].view(-1)[5] = 1
    clf = [t.clone().requires_grad_(True) for t in p["clf"]]
    _, cl = env.nt.det_clf_loss(p["ro"], clf, p["do"], p["det_gt"], mask, lab, [t.clone() for t in p["iou"]])
    assert torch.isnan(cl)
    cl.backward()
    assert torch.isnan(clf[1].grad.view(-1, 11)[5]).all() and not torch.isnan(clf[0].grad).any()
