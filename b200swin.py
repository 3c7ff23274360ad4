"""Import shim: ``import b200swin`` loads the package that lives in
``multi-modal-monodepth-estimation_b200/`` (a directory name Python cannot import directly)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi-modal-monodepth-estimation_b200")
_spec = importlib.util.spec_from_file_location("b200swin", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200swin"] = _mod
_spec.loader.exec_module(_mod)
