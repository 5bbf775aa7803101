"""Import alias: the package directory carries the repository's (hyphenated) name, which is not a
Python identifier, so `import egn_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "edge-guided-near-eye-image-analysis-for-head-mounted-displays_b200")
_spec = importlib.util.spec_from_file_location("egn_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["egn_b200"] = _mod
_spec.loader.exec_module(_mod)
