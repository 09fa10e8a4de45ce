"""Import alias for the product package.

The package directory is `road-object-detection-for-bdd100k_b200/` (the name the
build contract fixes); hyphens are not importable, so `rodet_b200` points its
`__path__` there and executes that directory's `__init__.py` in this namespace.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "road-object-detection-for-bdd100k_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
