"""rodet_b200 — B200-native box-level hot path (anchors, ARM/ODM matching + encoding,
decode, select, top-k, NMS) behind the function signatures of
YoungYoung619/road-object-detection-for-bdd100k (`utils/net_tools.py`,
`utils/common_tools.py`, `utils/tf_extended/bboxes.py`).

Host code is Python; tensors are CUDA `torch.Tensor`s handed zero-copy (DLPack) through a
ctypes C ABI (`include/rodet_b200.h`) into hand-written sm_100a kernels
(`csrc/*.cu` -> `librodet_b200.so`).  There is no CPU fallback: importing the package
without the built library raises.
"""
__version__ = "0.1.0"

from . import _abi            # noqa: F401  (loads librodet_b200.so or raises)
from . import config          # noqa: F401
from .anchor_table import AnchorTable   # noqa: F401
