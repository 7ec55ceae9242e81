"""Import the UNMODIFIED reference modules in the build container (TEST INFRASTRUCTURE; golden generation only).

* ``oracle/_shim`` supplies gpytorch / linear_operator / entmax / ftfy (absent here, see its README);
* ``/root/reference`` goes first on sys.path so that its ``utils`` / ``datasets`` / ``clip`` packages win;
* ``trainers/__init__.py`` is bypassed (it imports a non-existent ``adapter_taskres``) by
  registering an empty package object whose ``__path__`` points at the reference directory.
"""
import importlib
import os
import sys
import types

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SHIM = os.path.join(ROOT, "oracle", "_shim")


def setup():
    if not os.path.isdir(REF):
        raise RuntimeError("the reference tree is only available in the build container")
    for p in (SHIM, REF):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    if "trainers" not in sys.modules:
        pkg = types.ModuleType("trainers")
        pkg.__path__ = [os.path.join(REF, "trainers")]
        sys.modules["trainers"] = pkg


def ref_module(name: str):
    """e.g. ref_module("trainers.gp_template_weigher")."""
    setup()
    return importlib.import_module(name)
