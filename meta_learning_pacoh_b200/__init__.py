"""B200-native engine for PACOH's meta-training hot path.

``meta_learning_pacoh_b200.meta_learn`` mirrors the reference's public ``meta_learn`` API;
``meta_learning_pacoh_b200.engine`` is the host layer over the C ABI (include/pacoh_b200.h) of libpacoh_b200.so.
Importing the package loads the shared library and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when libpacoh_b200.so is missing)
from . import engine  # noqa: F401

__version__ = "0.1.0"
