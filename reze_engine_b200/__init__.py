"""Import shim: the product package lives in ``reze-engine_b200/`` (the name the
layout contract asks for); a hyphen is not importable, so this module re-points
``reze_engine_b200`` at that directory and executes its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "reze-engine_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
