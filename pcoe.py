"""Importable alias: the package directory name (``3d-pointcloud-orientation-estimation_b200``) is
not a Python identifier, so ``import pcoe`` loads it through importlib and aliases every submodule."""
import importlib
import sys

_REAL = "3d-pointcloud-orientation-estimation_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["pcoe" + _name[len(_REAL):]] = _mod
