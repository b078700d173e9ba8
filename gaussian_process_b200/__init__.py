"""gaussian_process_b200 -- B200-native (sm_100a) exact Gaussian-process engine.

Drop-in modules (same names / functions as happyjin/Gaussian_process): ``GP_regression``,
``tune_hyperparms_regression``, ``CO2_example``, ``GP_binary_classification``,
``GP_multi_classification``.  They are imported lazily so that ``import gaussian_process_b200`` works
on a machine without a GPU (the C-ABI library is only loaded when an engine is created).
"""
from ._lib import GpxError, LIB_PATH, load as load_library  # noqa: F401
from .engine import Engine, GPFit, get_engine, padded  # noqa: F401

__all__ = ["Engine", "GPFit", "get_engine", "padded", "GpxError", "load_library", "LIB_PATH"]
__version__ = "0.1.0"
