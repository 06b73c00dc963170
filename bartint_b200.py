"""Import alias: the package directory is named ``bart-int_b200`` (the layout the project brief asks
for), which is not a valid Python identifier.  This module makes it importable as ``bartint_b200``:
it turns itself into a package whose search path is that directory and runs its ``__init__``."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "bart-int_b200")
__path__ = [_pkg_dir]
__package__ = __name__
__file__ = _os.path.join(_pkg_dir, "__init__.py")
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f
