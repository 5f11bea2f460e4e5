"""Import alias: `import gnn_b200` loads the package directory `graph-neural-networks_b200/`
(whose name, fixed by the project layout, is not a valid Python identifier)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph-neural-networks_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
