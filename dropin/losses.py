"""Drop-in shim: put this directory in front of the reference's ``src/`` on ``sys.path`` and
``import losses`` resolves to the B200 implementation."""
from titanet_b200.losses import *  # noqa: F401,F403
from titanet_b200 import losses as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
