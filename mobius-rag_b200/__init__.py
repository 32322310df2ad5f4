"""mobius-rag_b200 -- B200-native exact cosine top-k scan behind Mobius-RAG's retrieval interface.

The directory name carries a hyphen (it is the project name); import it as ``mrag_b200`` (the
alias module at the repository root) or with ``importlib.import_module("mobius-rag_b200")``.

Layout
  csrc/            CUDA kernels (sm_100a) + the C ABI of include/mrag.h  -> libmrag.so
  _native.py       ctypes binding of that ABI (fails loudly if the library is missing)
  index.py         Index / Filter: one row shard in HBM
  table.py         PublishedTable: host half of rag_published_embeddings (ids, text, vocabularies)
  columns.py       the columnar buffers behind it (numpy; snapshot = np.savez, no pickle)
  multi.py         MultiIndex: several row shards of one table owned by one process (all GPUs of a box behind one store)
  vector_store.py  VectorStore ABC + B200VectorStore + get_vector_store()  (reference: vector_store.py)
  corpus_search.py vector_arm / _vector_arm                                (reference: corpus_search.py:1427)
  hybrid.py        hybrid rerank fused with the scan (reference: corpus_search.py:1909-2297)
  sharded.py       row-sharded search across GPUs (allgather + k-way merge)
  synth.py         deterministic synthetic corpora / metadata / queries
"""
from .vector_store import B200ChromaVectorStore, B200VectorStore, NoopVectorStore, VectorStore, get_vector_store  # noqa: F401
from .corpus_search import CorpusFilters, LexiconExpansion, _vector_arm, vector_arm  # noqa: F401
from .table import PublishedTable  # noqa: F401
from .index import Filter, Index, make_meta, merge_topk  # noqa: F401
from .multi import MultiIndex  # noqa: F401
from .hybrid import HybridTable, dtag_arm, hybrid_rerank, rrf_merge  # noqa: F401

__version__ = "0.1.0"
