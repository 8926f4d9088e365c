"""Import shim: makes the hyphenated package directory ``vit-grid-model_b200/`` importable as
``vit_grid_model_b200`` (a module that sets ``__path__`` is a package)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vit-grid-model_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
