"""Drop-in for /root/reference/model/loss/rnnt_loss.py."""
from speech2text_b200.loss.rnnt_loss import RnntLoss, RnntLossConfig  # noqa: F401
