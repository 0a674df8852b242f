"""Drop-in shim: put this directory in front of the reference's ``src/`` on ``sys.path`` and
``import transforms`` resolves to the B200 implementation."""
from titanet_b200.transforms import *  # noqa: F401,F403
from titanet_b200 import transforms as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
