"""Importable alias of the product package directory ``haplohyped-varawareml_b200/``.

The directory name required by the project layout contains a hyphen, which Python cannot import;
this shim points the package search path at it so that
``import haplohyped_varawareml_b200.capi`` etc. resolve to the files that live there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "haplohyped-varawareml_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
