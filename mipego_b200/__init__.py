"""mipego_b200 -- the importable name of the package whose sources live in `mip-ego_b200/`.

The repository layout prescribes the directory name `mip-ego_b200/`, which is not a Python
identifier.  This regular package therefore adds that directory to its own module search path
(`__path__`, the mechanism `pkgutil.extend_path` and namespace packages use), so that
`mipego_b200.gpr`, `mipego_b200._cabi`, ... resolve to the files there, and re-exports the public
names from `mipego_b200._api`.  No code is executed from strings; `importlib.reload` and tooling
see ordinary modules.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mip-ego_b200"))

from ._api import *  # noqa: E402,F401,F403
from ._api import __version__, _cabi, dist  # noqa: E402,F401
