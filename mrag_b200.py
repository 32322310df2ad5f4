"""Import alias: the package directory is ``mobius-rag_b200`` (hyphenated project name), which the
``import`` statement cannot spell.  ``import mrag_b200`` returns that package, and
``mrag_b200.<submodule>`` is the very same module object as ``mobius-rag_b200.<submodule>``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_real = "mobius-rag_b200"
_pkg = importlib.import_module(_real)
for _sub in ("_native", "build", "index", "vocab", "columns", "multi", "table", "vector_store", "corpus_search", "hybrid", "sharded", "synth"):
    sys.modules[__name__ + "." + _sub] = importlib.import_module(_real + "." + _sub)
sys.modules[__name__] = _pkg
