"""Import alias: `import wap_b200` == the package in ./watch-attend-and-parse-tensorflow-version_b200/
(whose directory name, inherited from the reference repository, is not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("watch-attend-and-parse-tensorflow-version_b200")
sys.modules[__name__] = _pkg
