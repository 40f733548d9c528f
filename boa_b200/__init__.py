"""Import shim: the package sources live in `body-and-organ-analysis_b200/` (a directory name Python cannot import
directly); `import boa_b200` resolves to them."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "body-and-organ-analysis_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
