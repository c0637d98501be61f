"""Mirror of the reference's ``model.loss`` package: only the hot-path modules live here; everything else
(``model.loss.cross_entropy`` etc.) still resolves to the reference checkout when it is on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
