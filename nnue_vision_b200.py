"""Import alias.  The package directory is `nnue-vision_b200/` (the repo-layout contract names it
with a hyphen, which Python cannot import); this shim loads it under the importable name
`nnue_vision_b200` so that `from nnue_vision_b200 import nnue, serialize, engine` works."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nnue-vision_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
