"""Import alias: the package directory is `gps-gaussian_b200/` (not a valid Python identifier),
so `import gps_gaussian_b200` resolves here and re-points the package path at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gps-gaussian_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
