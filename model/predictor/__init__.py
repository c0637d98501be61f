"""Mirror of the reference's ``model.predictor`` package: only the stateless predictor lives here;
``model.predictor.predictor`` (the factory) and the LSTM predictor still resolve to the reference checkout."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
