"""Import shim: the product package directory is ``ddpg-trucktrailer_b200/`` (the name the project layout
prescribes), which is not a legal Python identifier.  ``import ddpg_trucktrailer_b200`` resolves to it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ddpg-trucktrailer_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _f
