from .box_projection_loss import BoxProjectionLoss, mil_loss, projection_losses  # noqa: F401
from .levelset_loss import (LCM, LevelsetLoss, LocalConsistencyModule, length_regularization, levelset_assembly,  # noqa: F401
                            region_levelset)
