"""acetn_b200 -- B200 (sm_100a) implementation of Ace-TN's CTMRG hot path behind the reference's entry points.

B200_AVAILABLE mirrors the reference's CUTENSOR_AVAILABLE flag (acetn/evolution/_extensions/__init__.py:9,25), but
requesting the backend without the library or a device raises instead of falling back (there is no CPU path)."""
from . import _lib

B200_AVAILABLE = _lib.available()

__all__ = ["B200_AVAILABLE"]
