"""Importable alias of the package directory
``depth-fusion-in-transformer-based-video-object-detection_b200/`` (hyphens are not legal in a
Python module name).  ``import dfvod_b200`` executes that directory's ``__init__.py`` as the
package ``dfvod_b200``; submodules resolve normally (``dfvod_b200.ops.modules`` ...)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "depth-fusion-in-transformer-based-video-object-detection_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _module
_spec.loader.exec_module(_module)
