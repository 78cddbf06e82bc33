"""Importable alias of ``radar-signal-simulation-and-target-detection_b200/`` (a hyphen cannot
appear in a Python import).  The alias package's search path *is* that directory, so
``rsp_b200.frame`` etc. are the files that live there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "radar-signal-simulation-and-target-detection_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
