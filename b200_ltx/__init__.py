"""b200_ltx — importable name of the package stored in `video-generation-for-human-avatars_b200/`
(a directory name Python cannot import directly).  This shim only extends the package search path;
all code lives in that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "video-generation-for-human-avatars_b200")
__path__.append(_PKG_DIR)

from ._version import __version__  # noqa: E402,F401
