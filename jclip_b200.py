"""Import shim: the product package lives in `jittor-clip-fewshot_b200/` (a name Python's import
statement cannot spell).  `import jclip_b200` loads it and aliases it -- and every submodule, so that
`from jclip_b200.runtime import get_context` yields the SAME module object as `jclip_b200.runtime` and
never a second copy with its own ctypes classes and contexts -- under this name."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "jittor-clip-fewshot_b200"
_ALIAS = __name__

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)

    def exec_module(self, module):      # already executed under its real name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    """`jclip_b200.x.y` -> the module object of `jittor-clip-fewshot_b200.x.y`."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), origin=real_spec.origin,
                                               is_package=real_spec.submodule_search_locations is not None)
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
sys.modules[_ALIAS] = _pkg
