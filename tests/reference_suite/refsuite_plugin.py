"""pytest plugin that runs the REFERENCE's own test files (unmodified, staged under baseline/_ref/test by
__graft_entry__.build()) against this repo's drop-in package.

Loaded with `-p refsuite_plugin` by tests/test_reference_suite.py.  Before collection it
  * makes `test` the reference's test package (the standard library has a package of that name),
  * and, unless REFSUITE_TARGET=reference, answers every `import draughts[.x]` with `draughts_b200[.x]`.
With REFSUITE_TARGET=reference nothing is aliased and the files test the reference itself (the harness check)."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
REF = ROOT / "baseline" / "_ref"


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name == "draughts" or name.startswith("draughts."):
            return importlib.machinery.ModuleSpec(name, self)
        return None

    def create_module(self, spec):
        return importlib.import_module("draughts_b200" + spec.name[len("draughts"):])

    def exec_module(self, module):
        pass


def _install():
    pkg = types.ModuleType("test")
    pkg.__path__ = [str(REF / "test")]
    sys.modules["test"] = pkg
    if os.environ.get("REFSUITE_TARGET", "shim") == "reference":
        sys.path.insert(0, str(REF))
        return
    sys.path.insert(0, str(ROOT / "py-draughts_b200"))
    sys.meta_path.insert(0, _Alias())


_install()
