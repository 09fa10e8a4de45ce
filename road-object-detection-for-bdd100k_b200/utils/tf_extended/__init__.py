"""Mirror of `utils/tf_extended` (imported as `tfe` by the reference scripts) for the
box-level hot path: bboxes, tensors, math helpers and the evaluation metrics."""
from .tensors import *   # noqa: F401,F403
from .bboxes import *    # noqa: F401,F403
from .math import *      # noqa: F401,F403
from .metrics import *   # noqa: F401,F403
