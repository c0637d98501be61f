"""Drop-in for /root/reference/model/loss/pruned_rnnt_loss.py."""
from speech2text_b200.loss.pruned_rnnt_loss import PrunedRnntLoss, PrunedRnntLossConfig  # noqa: F401
