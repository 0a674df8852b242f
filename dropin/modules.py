"""Drop-in shim: put this directory in front of the reference's ``src/`` on ``sys.path`` and
``import modules`` resolves to the B200 implementation."""
from titanet_b200.modules import *  # noqa: F401,F403
from titanet_b200 import modules as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
