"""Import shim: the package directory is named ``flux-fp8-api_b200`` (not a valid Python identifier),
so this module turns itself into that package: ``import flux_fp8_api_b200`` (+ submodules) works from
the repo root without installing anything."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "flux-fp8-api_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _os
