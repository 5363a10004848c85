"""Import shim: the package directory is named `dcl-net_b200/` (not a valid Python
identifier), so `import dcl_net_b200` resolves here and re-points the package at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dcl-net_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
