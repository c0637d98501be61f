"""Import-path mirror of the reference's ``model`` package for the hot path only:
put /root/repo ahead of the reference on sys.path and
``from model.joiner.joiner import Joiner`` (rnnt_task.py:28-29) resolves here."""
from pkgutil import extend_path

# let ``model.encoder`` etc. still resolve to the reference checkout when it is on sys.path
__path__ = extend_path(__path__, __name__)
