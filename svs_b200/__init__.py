"""svs_b200 -- a B200-native (sm_100a) engine for the one data-parallel hot path of Rhobota/svs:
the query-vs-all-documents similarity and top-n selection behind `svs.KB.retrieve` /
`svs.AsyncKB.retrieve`.

    Engine                  ctypes face of the C ABI in include/svsb200.h (libsvsb200.so)
    DeviceEmbeddingsMatrix  drop-in for svs.kb._EmbeddingsMatrix (get_sync / get / invalidate)
    install(svs)            patch an imported `svs` package so that KB / AsyncKB use the engine

There is no CPU fallback: importing works anywhere, but creating an Engine without the built
library or without a B200 raises.
"""
from ._lib import EngineError, LIB_PATH
from .engine import Engine, Snapshot, launch_count, pinned_empty
from .matrix import DeviceEmbeddingsMatrix, DeviceMatrix, load_from_connection
from .dropin import install, uninstall

__version__ = "0.1.0"

__all__ = [
    "Engine", "Snapshot", "EngineError", "DeviceEmbeddingsMatrix", "DeviceMatrix",
    "load_from_connection", "install", "uninstall", "launch_count", "pinned_empty", "LIB_PATH",
]
