"""Import alias: the product package lives in the directory `smart-calibration_b200/` (the name the build
contract fixes), which is not a valid identifier.  `import smartcal_b200` resolves to that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smart-calibration_b200")
_spec = importlib.util.spec_from_file_location("smartcal_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["smartcal_b200"] = _mod
_spec.loader.exec_module(_mod)
