"""tes_b200 -- B200 (sm_100a) implementation of the Trusted-Maximizers Entropy Search acquisition
hot path, behind the reference's own function names.

The directory name carries hyphens (it is fixed by the project layout); import it as `tes_b200`
through the shim module at the repository root.
"""
from . import _lib  # noqa: F401
from ._lib import TesError, build  # noqa: F401

__all__ = ["TesError", "build"]
