"""Drop-in for /root/reference/model/loss/ctc_loss.py."""
from speech2text_b200.loss.ctc_loss import CtcLoss, CtcLossConfig  # noqa: F401
