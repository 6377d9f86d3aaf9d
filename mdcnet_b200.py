"""Importable alias: the package directory keeps the repository's mandated name
`mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200/` (not a valid Python
identifier); `import mdcnet_b200` loads it under this name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "mdc-net-multimodal-defect-captioning-network-for-surface-steel-defects_b200")
_spec = importlib.util.spec_from_file_location("mdcnet_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mdcnet_b200"] = _mod
_spec.loader.exec_module(_mod)
