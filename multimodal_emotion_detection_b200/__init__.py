"""Importable alias of the package directory ``multimodal-emotion-detection_b200/`` (a hyphen cannot appear in
a Python module name).  All code lives there; this file only points the import system at it."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multimodal-emotion-detection_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
