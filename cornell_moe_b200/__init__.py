"""Importable alias for the package directory ``cornell-moe_b200/`` (a hyphen is not a legal module name).

All code lives in ``cornell-moe_b200/``; this stub only points Python's import machinery at it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cornell-moe_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
