"""Import shim: the package directory carries the reference's repo name
(`cse375-finalproj-huffman-decoding_b200`), which is not a valid Python
identifier, so it is loaded by path and re-exported under `hd_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cse375-finalproj-huffman-decoding_b200")
_NAME = "cse375_finalproj_huffman_decoding_b200"

if _NAME in sys.modules:
    _mod = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
PKG_DIR = _PKG_DIR
