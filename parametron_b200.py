"""Import shim: the package directory is `parametron.jl_b200/` (a dot in the name), which `import` cannot spell.
`import parametron_b200` loads that directory as the package `parametron_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parametron.jl_b200")
_spec = importlib.util.spec_from_file_location("parametron_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["parametron_b200"] = _mod
_spec.loader.exec_module(_mod)
