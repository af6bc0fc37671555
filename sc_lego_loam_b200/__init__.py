"""Import alias: the product package lives in the directory ``sc-lego-loam_b200/`` (the name the build
contract fixes); a hyphen is not importable, so ``import sc_lego_loam_b200`` resolves here and this shim
points the package at the real directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sc-lego-loam_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
