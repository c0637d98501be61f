"""speech2text_b200: sm_100a implementation of guangkun0818/speech2text's
transducer-loss hot path (pruned / full RNN-T loss + joiner), behind the
reference's own ``model.joiner`` / ``model.loss`` module API."""
from .joiner import Joiner, JoinerConfig, LazyJoinerLogits  # noqa: F401
from .loss import Loss, PrunedRnntLoss, PrunedRnntLossConfig, RnntLoss, RnntLossConfig  # noqa: F401
from .predictor import StatelessPredictor, StatelessPredictorConfig  # noqa: F401

__version__ = "0.1.0"
