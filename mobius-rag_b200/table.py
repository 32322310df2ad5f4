"""PublishedTable: the host-side half of ``rag_published_embeddings`` + ``document_tags``.

The GPU index (Index) holds the vectors and the coded filter columns; this object holds what
the reference hydrates from the same SELECT (`_BM25_COLS`, corpus_search.py:621-640): ids,
text, page numbers, tags -- plus the string->code vocabularies and the document_id -> doc_idx map.
"""
from __future__ import annotations

import threading
from typing import Any, Sequence

import numpy as np

from . import _native as N
from .index import Filter, Index, make_meta
from .vocab import Vocab

# corpus_search.py:208-213
FL_MEDICAID_MCO_PAYERS = frozenset({
    "Sunshine Health", "Simply Healthcare", "United Healthcare",
    "Aetna", "Molina Healthcare", "Molina Healthcare of Florida",
    "WellCare", "Humana", "Humana Healthy Horizons",
})
FL_STATE_AUTHORITY_PAYERS = ["AHCA", "Ahca.myflorida", "Florida Medicaid"]

HYDRATE_COLS = (
    "text", "page_number", "paragraph_index", "section_path", "chapter_path", "summary", "content_sha",
    "document_display_name", "document_filename", "chunk_d_tags", "chunk_p_tags", "chunk_j_tags",
)


def to_float4(emb: Sequence[float]) -> np.ndarray:
    """The stored / queried numeric format: text repr(float(x)) parsed by pgvector's strtof
    (embedding_worker.py:53-62, vector_store.py:272) == np.float32(python float)."""
    return np.asarray([float(x) for x in emb], dtype=np.float64).astype(np.float32)


class PublishedTable:
    def __init__(self, dim: int, dtype: str = "f32", device: int = 0, capacity: int = 1 << 20):
        self.index = Index(dim, dtype, device, capacity)
        self.vocab = Vocab()
        self.lock = threading.RLock()
        self.id: list[str] = []
        self.document_id: list[str] = []
        self.source_type: list[str | None] = []
        self.source_id: list[str | None] = []
        self.document_payer: list[str | None] = []
        self.document_state: list[str | None] = []
        self.document_program: list[str | None] = []
        self.document_authority_level: list[str | None] = []
        self.extra: dict[str, list] = {c: [] for c in HYDRATE_COLS}
        self.doc_idx: dict[str, int] = {}
        self._next_doc = 0          # doc_idx values are never reused (tombstoned rows keep theirs)
        self.doc_d_tags: dict[str, set] = {}
        self.doc_p_tags: dict[str, set] = {}

    def __len__(self) -> int:
        return len(self.id)

    # -- write side (publish.py:204-362 / embedding_worker.py:229-266) -------------------------
    def _doc(self, document_id: str) -> int:
        d = self.doc_idx.get(document_id)
        if d is None:
            d = self._next_doc
            self._next_doc += 1
            self.doc_idx[document_id] = d
        return d

    def insert(self, rows: Sequence[dict[str, Any]], embeddings: Sequence[Sequence[float] | None]) -> None:
        """INSERT rows; ``embeddings[i] is None`` leaves embedding_vec NULL for that row."""
        n = len(rows)
        if n == 0:
            return
        if len(embeddings) != n:
            raise ValueError("rows / embeddings length mismatch")
        dim = self.index.dim
        X = np.zeros((n, dim), dtype=np.float32)
        valid = np.ones(n, dtype=np.uint8)
        for i, e in enumerate(embeddings):
            if e is None:
                valid[i] = 0
                continue
            v = to_float4(e)
            if v.shape[0] != dim:
                raise ValueError(f"expected {dim} dimensions, not {v.shape[0]}")   # pgvector's error text
            X[i] = v
        with self.lock:
            v = self.vocab
            meta = make_meta(
                n,
                doc_idx=[self._doc(str(r["document_id"])) for r in rows],
                payer=[v.payer.encode(r.get("document_payer")) for r in rows],
                state=[v.state.encode(r.get("document_state")) for r in rows],
                program=[v.program.encode(r.get("document_program")) for r in rows],
                authority=[v.authority.encode(r.get("document_authority_level")) for r in rows],
                source_type=[v.source_type.encode(r.get("source_type")) for r in rows],
                valid=valid,
            )
            first = self.index.append(X, meta)
            assert first == len(self.id), "host table and device index out of step"
            for r in rows:
                self.id.append(str(r["id"]))
                self.document_id.append(str(r["document_id"]))
                self.source_type.append(r.get("source_type"))
                self.source_id.append(None if r.get("source_id") is None else str(r.get("source_id")))
                self.document_payer.append(r.get("document_payer"))
                self.document_state.append(r.get("document_state"))
                self.document_program.append(r.get("document_program"))
                self.document_authority_level.append(r.get("document_authority_level"))
                for c in HYDRATE_COLS:
                    self.extra[c].append(r.get(c))

    def set_document_tags(self, document_id: str, d_tags: Sequence[str] | None, p_tags: Sequence[str] | None) -> None:
        """UPSERT one document_tags row (keys of d_tags / p_tags, app/models.py:525-543)."""
        with self.lock:
            d = self._doc(str(document_id))
            self.doc_d_tags[str(document_id)] = set(d_tags or ())
            self.doc_p_tags[str(document_id)] = set(p_tags or ())
            bits = np.zeros((1, N.MRAG_TAG_WORDS), dtype=np.uint64)
            for kind, keys in (("d", d_tags or ()), ("p", p_tags or ())):
                for key in keys:
                    b = self.vocab.tag_bit(kind, key, allocate=True)
                    bits[0, b >> 6] |= np.uint64(1 << (b & 63))
            self.index.set_doc_tags(d, bits)

    def delete_document(self, document_id: str) -> int:
        """DELETE FROM .. WHERE document_id = :id (publish.py:310-313)."""
        with self.lock:
            d = self.doc_idx.get(str(document_id))
            if d is None:
                return 0
            n = self.index.tombstone_doc(d)
            # a re-published document gets a fresh doc_idx so the tombstoned rows stay dead
            del self.doc_idx[str(document_id)]
            self.doc_d_tags.pop(str(document_id), None)
            self.doc_p_tags.pop(str(document_id), None)
            return n

    # -- snapshot ------------------------------------------------------------------------------
    _HOST_STATE = ("id", "document_id", "source_type", "source_id", "document_payer", "document_state", "document_program",
                   "document_authority_level", "extra", "doc_idx", "_next_doc", "doc_d_tags", "doc_p_tags")

    def save(self, dirpath: str, corpus_version: int = 0) -> None:
        """Snapshot = the device shard (``index.mrag``) + the host half of the table (``table.pkl``): ids, hydration
        columns, vocabularies, document maps.  Keyed by corpus_state.corpus_version (publish.py:314)."""
        import os
        import pickle
        os.makedirs(dirpath, exist_ok=True)
        with self.lock:
            self.index.save(os.path.join(dirpath, "index.mrag"), corpus_version)
            state = {k: getattr(self, k) for k in self._HOST_STATE}
            state["vocab"] = self.vocab
            state["corpus_version"] = int(corpus_version)
            with open(os.path.join(dirpath, "table.pkl"), "wb") as f:
                pickle.dump(state, f, protocol=pickle.HIGHEST_PROTOCOL)

    @classmethod
    def load(cls, dirpath: str, device: int = 0, capacity: int = 0) -> tuple["PublishedTable", int]:
        """(table, corpus_version) from a snapshot directory; the cold start is one sequential read."""
        import os
        import pickle
        with open(os.path.join(dirpath, "table.pkl"), "rb") as f:
            state = pickle.load(f)
        self = cls.__new__(cls)
        self.index, ver = Index.load(os.path.join(dirpath, "index.mrag"), device, capacity)
        if ver != state["corpus_version"] or len(self.index) != len(state["id"]):
            self.index.close()
            raise ValueError("snapshot is inconsistent: index.mrag and table.pkl come from different versions")
        self.lock = threading.RLock()
        self.vocab = state["vocab"]
        for k in cls._HOST_STATE:
            setattr(self, k, state[k])
        return self, ver

    # -- WHERE builders ------------------------------------------------------------------------
    def filter_pg_store(self, document_id: str | None, filters: dict | None) -> Filter | None:
        """WHERE of PgVectorStore._search_async (vector_store.py:245-267)."""
        f = Filter()
        v = self.vocab
        if document_id:
            d = self.doc_idx.get(str(document_id))
            f.doc_eq(0xFFFFFFFF if d is None else d)
        for key, value in (filters or {}).items():
            if value is None or value == "":
                continue
            if key == "payer":
                f.payer_in([v.payer.lookup(value)])
            elif key == "state":
                f.state_eq(v.state.lookup(value))
            elif key == "authority_level":
                f.authority_eq(v.authority.lookup(value))
            elif key == "document_id":
                d = self.doc_idx.get(str(value))
                if f.s.flags & N.F_DOC_EQ and f.s.doc_eq != (0xFFFFFFFF if d is None else d):
                    f.doc_eq(0xFFFFFFFF)        # document_id = a AND document_id = b, a != b
                else:
                    f.doc_eq(0xFFFFFFFF if d is None else d)
            elif key == "source_type":
                f.source_type_eq(v.source_type.lookup(value))
            # unknown keys: skipped silently (vector_store.py:253-258)
        return f if f.active else None

    def filter_corpus(self, filters: Any, include_document_ids: Sequence[str] | None) -> Filter:
        """_build_filter_clauses (corpus_search.py:516-560)."""
        f = Filter()
        v = self.vocab
        if filters:
            payer = getattr(filters, "payer", None)
            if payer:
                if payer in FL_MEDICAID_MCO_PAYERS:
                    f.payer_in([v.payer.lookup(payer)],
                               [v.payer.lookup(p) for p in FL_STATE_AUTHORITY_PAYERS], v.state.lookup("FL"))
                else:
                    f.payer_in([v.payer.lookup(payer)])
            if getattr(filters, "state", None):
                f.state_eq(v.state.lookup(filters.state))
            if getattr(filters, "program", None):
                f.program_eq(v.program.lookup(filters.program))
            if getattr(filters, "authority_level", None):
                f.authority_eq(v.authority.lookup(filters.authority_level))
        if include_document_ids:
            f.doc_pool([d for d in (self.doc_idx.get(str(x)) for x in include_document_ids) if d is not None])
        return f
