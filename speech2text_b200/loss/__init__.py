from .loss import Loss  # noqa: F401
from .pruned_rnnt_loss import PrunedRnntLoss, PrunedRnntLossConfig  # noqa: F401
from .rnnt_loss import RnntLoss, RnntLossConfig  # noqa: F401
