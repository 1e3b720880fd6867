from .assigner import AssignResult, ClassificationCost, MaskHungarianAssigner  # noqa: F401
from .match_cost import BoxMatchingCost, projection_profiles  # noqa: F401
