"""ctypes binding of libmhmr_sm100.so (the C-ABI declared in include/mhmr.h).

The library is built in-tree by build.py.  Importing this module never falls back to another
implementation: if the shared object is missing it is (re)built, and if that fails the import raises.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmhmr_sm100.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mhmr.h")

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

_lib = None


class MhmrError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Function names declared in include/mhmr.h (used by the export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mhmr_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        import importlib.util

        spec = importlib.util.spec_from_file_location("_mhmr_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(LIB_PATH)
    lib.mhmr_last_error.restype = ctypes.c_char_p
    lib.mhmr_last_error.argtypes = []
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mhmr_last_error().decode("utf-8", "replace")
        exc = MhmrError if rc != -2 else AssertionError
        raise exc(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> c_void_p:
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
