"""Import shim: `import litejpeg_b200` loads the package that lives in the
directory `jpeg_encoder_gsoc2017_ishan-bansal_b200/` (a '-' cannot appear in an
import statement) under the module name `litejpeg_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jpeg_encoder_gsoc2017_ishan-bansal_b200")
_spec = importlib.util.spec_from_file_location("litejpeg_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["litejpeg_b200"] = _mod
_spec.loader.exec_module(_mod)
