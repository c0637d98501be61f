"""Drop-in for /root/reference/model/predictor/stateless_predictor.py."""
from speech2text_b200.predictor import StatelessPredictor, StatelessPredictorConfig  # noqa: F401
