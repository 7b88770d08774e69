"""``import strainer_b200`` -> the package in ./strainer-gan_b200/ (a directory name that is not a
valid identifier, kept because it is the layout this build was asked for)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "strainer-gan_b200")
_name = "strainer_gan_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg_dir, "__init__.py"),
                                                   submodule_search_locations=[_pkg_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_name]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
__version__ = _mod.__version__
