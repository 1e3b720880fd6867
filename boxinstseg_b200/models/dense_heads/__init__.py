from .condinst_mask_head import CondInstMaskHead  # noqa: F401
