"""Drop-in shim: put this directory in front of the reference's ``src/`` on ``sys.path`` and
``import models`` resolves to the B200 implementation."""
from titanet_b200.models import *  # noqa: F401,F403
from titanet_b200 import models as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
import losses, modules  # noqa: E402,F401  (the reference's models.py imports these names too)
