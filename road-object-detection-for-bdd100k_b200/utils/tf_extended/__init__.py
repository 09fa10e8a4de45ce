"""Mirror of `utils/tf_extended` (imported as `tfe` by the reference scripts) for the
box-level hot path: bboxes, tensors and math helpers."""
from .tensors import *   # noqa: F401,F403
from .bboxes import *    # noqa: F401,F403
from .math import *      # noqa: F401,F403
