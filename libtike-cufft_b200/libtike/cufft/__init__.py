"""libtike.cufft -- B200-native drop-in for the ptychography hot path of libtike-cufft.

Same import surface as the reference package (src/libtike/cufft/__init__.py:1-9):
`from libtike.cufft.ptycho import *` plus `__version__`.  `libtike` itself stays a
namespace package (no libtike/__init__.py), as in the reference (setup.py:24).
"""
from libtike.cufft.ptycho import *  # noqa: F401,F403

try:
    from importlib.metadata import version, PackageNotFoundError
    try:
        __version__ = version("libtike-cufft")
    except PackageNotFoundError:  # package is not installed
        __version__ = "0.4.0+b200"
except ImportError:  # pragma: no cover
    __version__ = "0.4.0+b200"
