"""Head-output layout contract of the reference's network factory (nets/catch_net.py:276-363).

The CNN itself (backbones, deconvolution, merge, conv heads) is outside the box-level hot path
(DESIGN.md section 8); what the path depends on is the LAYOUT the heads hand over:
  refine_out / det_out: list per layer of [bs, fh, fw, A, 4]      (nets/catch_net.py:306-308)
  clf_out:              list per layer of [bs, fh, fw, A, n_obj]  (:339-341, logits, no softmax)
and the selection rule of get_output (:344-363).  These helpers turn the NHWC outputs of any
convolutional head (channels = A*4 or A*n_obj) into that layout as zero-copy views, so the tensors
can go straight into net_tools.det_groundtruth / decode_detected_bboxes."""
from __future__ import annotations

from .. import config
from ..utils import net_tools


def head_shapes(batch, backbone_name='mobilenet_v2', feat_sizes=None):
    """([bs,fh,fw,A,4] per layer, [bs,fh,fw,A,n_obj] per layer) for a backbone
    (config.feat_size_all_layers, valid for 418x418; pass feat_sizes for other input sizes)."""
    n_anchor = net_tools.n_anchor_each_layer(backbone_name)
    feats = feat_sizes if feat_sizes is not None else config.feat_size_all_layers[backbone_name]
    sizes = list(feats.values()) if isinstance(feats, dict) else list(feats)
    assert len(sizes) == len(n_anchor), "one feature-map size per extracted layer"
    det = [(batch, fh, fw, a, 4) for (fh, fw), a in zip(sizes, n_anchor)]
    clf = [(batch, fh, fw, a, config.total_obj_n) for (fh, fw), a in zip(sizes, n_anchor)]
    return det, clf


def _split_channels(conv_outputs, backbone_name, inner):
    n_anchor = net_tools.n_anchor_each_layer(backbone_name)
    outs = list(conv_outputs.values()) if isinstance(conv_outputs, dict) else list(conv_outputs)
    if len(outs) != len(n_anchor):
        raise ValueError("expected %d head tensors, got %d" % (len(n_anchor), len(outs)))
    res = []
    for t, a in zip(outs, n_anchor):
        if t.dim() != 4 or t.shape[-1] != a * inner:
            raise ValueError("head tensor must be NHWC with %d channels, got shape %s" % (a * inner, tuple(t.shape)))
        res.append(t.reshape(t.shape[0], t.shape[1], t.shape[2], a, inner))      # tf.reshape(output, [-1]+hw+[A,inner])
    return res


def det_out(conv_outputs, backbone_name='mobilenet_v2'):
    """[bs,fh,fw,A*4] per layer -> [bs,fh,fw,A,4] per layer (nets/catch_net.py:306-308); also for refine_out."""
    return _split_channels(conv_outputs, backbone_name, 4)


def clf_out(conv_outputs, backbone_name='mobilenet_v2'):
    """[bs,fh,fw,A*n_obj] per layer -> [bs,fh,fw,A,n_obj] logits per layer (nets/catch_net.py:339-341)."""
    return _split_channels(conv_outputs, backbone_name, config.total_obj_n)


class factory(object):
    """The output side of the reference's `factory` (nets/catch_net.py:37-96, 344-363) for heads computed
    elsewhere: holds refine_out / det_out / clf_out in the hot path's layout and applies get_output's
    selection rule.  `config_dict['train_range']` is config.train_range.ALL or .REFINE."""

    def __init__(self, refine_conv, det_conv=None, clf_conv=None, backbone_name='mobilenet_v2', is_training=False,
                 config_dict=None):
        assert backbone_name in config.supported_backbone_name
        self.backbone_name = backbone_name
        self.is_training = is_training
        self.train_range = (config_dict or {}).get('train_range', config.train_range.ALL)
        self.refine_out = det_out(refine_conv, backbone_name)
        if self.train_range is config.train_range.ALL:
            self.det_out = det_out(det_conv, backbone_name)
            self.clf_out = clf_out(clf_conv, backbone_name)

    def get_output(self):
        """ALL -> (refine_out, det_out, clf_out); REFINE -> refine_out; else ValueError('Error')."""
        if self.train_range is config.train_range.ALL:
            return self.refine_out, self.det_out, self.clf_out
        elif self.train_range is config.train_range.REFINE:
            return self.refine_out
        raise ValueError('Error')
