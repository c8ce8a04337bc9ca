"""Import alias: the product package lives in the directory `bind-your-avatar-implementation_b200/` (a name
Python cannot import directly because of the hyphens).  `import bya_b200` loads that directory as the package
`bya_b200`, with normal sub-module imports (`bya_b200.ops`, `bya_b200.transformer`, ...)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "bind-your-avatar-implementation_b200")
_spec = importlib.util.spec_from_file_location(
    "bya_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bya_b200"] = _mod
_spec.loader.exec_module(_mod)
