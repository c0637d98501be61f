"""Drop-in for /root/reference/model/loss/loss.py."""
from speech2text_b200.loss.loss import Loss  # noqa: F401
