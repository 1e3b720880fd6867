from .pairwise import pairwise_nlog  # noqa: F401  (same export as mmdet/ops/pairwise/__init__.py:1)
