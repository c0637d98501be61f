"""Drop-in for /root/reference/model/joiner/joiner.py: re-exports the sm_100a implementation."""
from speech2text_b200.joiner import Joiner, JoinerConfig, LazyJoinerLogits  # noqa: F401
