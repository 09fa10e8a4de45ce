"""Synthetic stand-ins for what the reference scripts get from TFRecords and from the CNN: ground-truth boxes
(rodet_b200.synth) and head outputs in the conv layout [bs, fh, fw, A * inner] that nets/catch_net.py reshapes
(the network itself is outside the box-level path)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rodet_b200 import config, synth                      # noqa: E402
from rodet_b200.nets.catch_net import factory             # noqa: E402
from rodet_b200.utils import net_tools                    # noqa: E402


def anchors_for(backbone_name="mobilenet_v2"):
    """train.py:95-99 / evaluate.py:91-95 / predict.py:67-71."""
    layer_n = len(list(config.extract_feat_name[backbone_name]))
    return net_tools.anchors_all_layer(config.img_size, config.feat_size_all_layers[backbone_name],
                                       net_tools.init_anchor(layer_n))


def ground_truth(first_image, batch, device):
    """(bboxes [B,G,4] corner form zero padded, labels [B,G] int64, counts [B] int32) on the device."""
    boxes, labels, counts = synth.gt_batch(first_image, batch)
    t = lambda a: torch.from_numpy(a).to(device)
    return t(boxes), t(labels), t(counts)


def fake_network(batch, device, seed, backbone_name="mobilenet_v2", train_range=None, refine_targets=None):
    """A `factory` whose heads are random conv outputs (NHWC, A * 4 / A * 11 channels per layer).  With
    `refine_targets` (per-layer [B,fh,fw,A,4] + mask lists) half of the positive anchors predict their ARM target
    up to noise, as a partly trained refine head would."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_anchor = net_tools.n_anchor_each_layer(backbone_name)
    feats = list(config.feat_size_all_layers[backbone_name].values())
    conv = lambda inner, sigma: [(torch.randn((batch, fh, fw, a * inner), generator=g) * sigma).to(device)
                                 for (fh, fw), a in zip(feats, n_anchor)]
    refine, det, clf = conv(4, 0.1), conv(4, 0.05), conv(config.total_obj_n, 2.0)
    if refine_targets is not None:
        gts, masks = refine_targets
        for r, gt, m, a in zip(refine, gts, masks, n_anchor):
            keep = (m.reshape(r.shape[0], r.shape[1], r.shape[2], a, 1) > 0)
            keep[..., ::2, :] = False
            r5 = r.view(r.shape[0], r.shape[1], r.shape[2], a, 4)
            r5.copy_(torch.where(keep, gt + 0.05 * r5, r5))
    for c in clf:
        c.view(*c.shape[:3], -1, config.total_obj_n)[..., 0] += 3.0      # background bias
    cd = {"train_range": train_range if train_range is not None else config.train_range.ALL}
    return factory(refine, det, clf, backbone_name=backbone_name, is_training=False, config_dict=cd)
