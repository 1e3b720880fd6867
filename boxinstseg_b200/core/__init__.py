from .match_cost import BoxMatchingCost, projection_profiles  # noqa: F401
