from .condinst import LogVars, mask_branch_step, parse_losses  # noqa: F401
