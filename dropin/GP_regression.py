"""Drop-in shim: put this directory on PYTHONPATH and `import GP_regression` resolves to the gpx B200 engine's
module of the same name (same functions / signatures as the reference script)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from gaussian_process_b200.GP_regression import *  # noqa: F401,F403,E402
from gaussian_process_b200 import GP_regression as _impl  # noqa: E402


def __getattr__(name):  # module globals the reference drivers read (true_fun, n, mu_post, ...)
    return getattr(_impl, name)
