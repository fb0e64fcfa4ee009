"""Import shim: the package directory is named `raytracing-in-windows-console_b200` (not a valid
Python identifier), so `import rtc_b200` loads it from there under this name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raytracing-in-windows-console_b200")
_spec = importlib.util.spec_from_file_location(
    "rtc_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rtc_b200"] = _mod
_spec.loader.exec_module(_mod)
