"""Import shim: the package directory is `rgbid-slam_b200/` (hyphen, as the layout contract names it), which is
not a valid Python identifier.  `import rgbid_slam_b200` loads that directory as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rgbid-slam_b200")
_spec = importlib.util.spec_from_file_location("rgbid_slam_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rgbid_slam_b200"] = _mod
_spec.loader.exec_module(_mod)
