"""Import alias: `import b200mm` loads the package that lives in `ant-multi-modal-framework_b200/`.

The directory name required by the project layout contains '-' and therefore is not importable by name; this shim
registers it in sys.modules under the name `b200mm` (sub-modules resolve normally: b200mm.ops, b200mm.modules.vit ...).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ant-multi-modal-framework_b200")
_spec = importlib.util.spec_from_file_location("b200mm", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200mm"] = _mod
_spec.loader.exec_module(_mod)
