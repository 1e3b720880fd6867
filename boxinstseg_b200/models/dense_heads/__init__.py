from .condinst_mask_head import CondInstMaskHead  # noqa: F401
from .meanfield import MeanField  # noqa: F401
