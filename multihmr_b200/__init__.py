"""Importable alias of the `multi-hmr_b200/` package directory (a hyphen is not a valid module name).

`import multihmr_b200.ops` resolves submodules from `multi-hmr_b200/`.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multi-hmr_b200")
__path__.insert(0, _real)
del _os, _real
