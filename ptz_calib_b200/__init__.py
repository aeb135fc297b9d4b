"""Import shim: the package directory is `ptz-calib_b200/` (not a valid Python identifier), so
`import ptz_calib_b200` resolves here and forwards to it."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ptz-calib_b200"))
from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: E402
