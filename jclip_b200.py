"""Import shim: the product package lives in `jittor-clip-fewshot_b200/` (a name Python's import
statement cannot spell).  `import jclip_b200` loads it and aliases it under this name."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("jittor-clip-fewshot_b200")
sys.modules[__name__] = _pkg
