"""Import alias: the package directory is ``mobius-rag_b200`` (hyphenated project name), which the
``import`` statement cannot spell.  ``import mrag_b200`` returns that package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("mobius-rag_b200")
sys.modules[__name__] = _pkg
