"""B200-native engine for PACOH's meta-training hot path.

``meta_learning_pacoh_b200.meta_learn`` mirrors the reference's public ``meta_learn`` API;
``meta_learning_pacoh_b200.engine`` is the host layer over the C ABI (include/pacoh_b200.h) of libpacoh_b200.so.
The shared library is loaded the first time ``_lib`` / ``engine`` / ``meta_learn`` is touched and that fails loudly
(ImportError) if it has not been built: there is no CPU fallback.  ``meta_learning_pacoh_b200.build`` (the nvcc
driver) is importable without the library, so a fresh tree can build itself.
"""
import importlib

__version__ = "0.1.0"
_LAZY = ("_lib", "engine", "meta_learn", "data_sim")


def __getattr__(name):
    if name in _LAZY:
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
