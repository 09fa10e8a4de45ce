"""Tier A: import the UNMODIFIED reference sources over the NumPy `tensorflow` shim.

TEST INFRASTRUCTURE ONLY.  Works only where `/root/reference` exists (the build
container); the GPU box uses the committed fixtures in `tests/golden/` instead.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RODET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "net_tools.py"))


def load_reference() -> types.SimpleNamespace:
    """Returns namespace(config, net_tools, common_tools, tfe, tf).

    The reference uses top-level module names `config` and `utils`; they are
    imported from REFERENCE_ROOT (which must come first on sys.path for this).
    """
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REFERENCE_ROOT)
    from oracle import tf_shim
    tf = tf_shim.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("config", "utils"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            raise RuntimeError("module %r already imported from elsewhere" % name)
    import config                      # /root/reference/config.py
    from utils import net_tools        # /root/reference/utils/net_tools.py
    from utils import common_tools
    import utils.tf_extended as tfe
    return types.SimpleNamespace(config=config, net_tools=net_tools,
                                 common_tools=common_tools, tfe=tfe, tf=tf)


def reference_anchors(ref, img_size, feat_sizes):
    """anchors_all_layer at an arbitrary (img_size, feat_sizes); `init_anchor`
    reads `config.img_size` globally (utils/net_tools.py:37-38)."""
    ref.config.img_size = tuple(img_size)
    feats = {"layer_%d" % (i + 1): tuple(fs) for i, fs in enumerate(feat_sizes)}
    return ref.net_tools.anchors_all_layer(tuple(img_size), feats,
                                           ref.net_tools.init_anchor(len(feat_sizes)))
