"""Importable alias for the package directory ``segment-anything-in-nerf_b200/``.

The product sources live in the hyphenated directory the build contract names;
a hyphen is not a legal Python identifier, so this stub extends ``__path__`` to
that directory and re-exports its public names.  Nothing else lives here.
"""
import os as _os

_PKG_DIR = _os.path.normpath(
    _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "segment-anything-in-nerf_b200")
)
__path__.insert(0, _PKG_DIR)  # type: ignore[name-defined]

from .config import GridConfig, SAMNeRFConfig  # noqa: E402,F401
from .synthetic import make_synthetic_params  # noqa: E402,F401
