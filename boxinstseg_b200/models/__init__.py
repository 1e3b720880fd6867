from .builder import HEADS, LOSSES, MATCH_COST, build_head, build_loss  # noqa: F401
from . import losses, dense_heads  # noqa: F401  (registers the classes)
from .detectors import mask_branch_step, parse_losses  # noqa: F401
