"""Mirror of the reference's ``model.joiner`` package (see model/__init__.py)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
