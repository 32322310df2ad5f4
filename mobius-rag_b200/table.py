"""PublishedTable: the host-side half of ``rag_published_embeddings`` + ``document_tags``.

The GPU index (Index, or MultiIndex when one process owns several shards) holds the vectors and the coded filter
columns; this object holds what the reference hydrates from the same SELECT (`_BM25_COLS`,
corpus_search.py:621-640): ids, text, page numbers, tags -- as COLUMNS (columns.py: numpy buffers, no per-row Python
objects, snapshot = ``np.savez``) -- plus the string->code vocabularies and the document_id -> doc_idx map.

Write paths (publish.py:204-362, embedding_worker.py:229-266):
  insert(rows, embeddings)        row dicts, as the worker's batches of 50 arrive
  insert_columns(X, columns)      bulk: an [n, dim] array + one sequence per column (a loader's path for 10M+ rows)
Rows become visible to searches only AFTER their host columns exist (the host columns are appended first, the device
rows second), so a concurrent search can never return a row it cannot hydrate.
"""
from __future__ import annotations

import json
import os
import threading
from typing import Any, Sequence

import numpy as np

from . import _native as N
from .columns import CodeCol, IntCol, JsonCol, StrCol
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
_INT_COLS = ("page_number", "paragraph_index")
_JSON_COLS = ("chunk_d_tags", "chunk_p_tags", "chunk_j_tags")
_CODED = (("source_type", "source_type", np.uint8), ("document_payer", "payer", np.uint16), ("document_state", "state", np.uint8),
          ("document_program", "program", np.uint8), ("document_authority_level", "authority", np.uint8))


def to_float4(emb: Sequence[float]) -> np.ndarray:
    """The stored / queried numeric format: text repr(float(x)) parsed by pgvector's strtof
    (embedding_worker.py:53-62, vector_store.py:272) == np.float32(python float)."""
    return np.asarray(emb, dtype=np.float64).astype(np.float32)


class _DocIdCol:
    """``document_id`` of a row = the id string of its document (rows store the dense doc index)."""

    def __init__(self, table: "PublishedTable"):
        self.t = table

    def __len__(self) -> int:
        return self.t._host_n

    def __getitem__(self, i: int) -> str:
        i = int(i)
        if i < 0:
            i += self.t._host_n
        if not 0 <= i < self.t._host_n:          # host columns are written BEFORE the device rows become searchable
            raise IndexError(i)
        return self.t.doc_ids[int(self.t.row_doc[i])]


class PublishedTable:
    def __init__(self, dim: int, dtype: str = "f32", device: int = 0, capacity: int = 1 << 20,
                 devices: Sequence[int] | None = None):
        """``devices``: GPUs that share the rows (one shard each; a device may be named twice).  None = one shard on
        ``device``."""
        if devices is not None and len(devices) > 1:
            from .multi import MultiIndex
            self.index = MultiIndex(dim, dtype, devices, capacity)
        else:
            self.index = Index(dim, dtype, device if devices is None else devices[0], capacity)
        self._init_host(Vocab())

    def _init_host(self, vocab: Vocab) -> None:
        self.vocab = vocab
        self.lock = threading.RLock()
        self._n = 0                                  # rows whose host columns AND device rows exist
        self._host_n = 0                             # rows whose host columns exist (>= _n while an insert is in flight)
        self.id = StrCol()
        self.source_id = StrCol()
        self.row_doc = np.zeros(1024, dtype=np.uint32)
        self.document_id = _DocIdCol(self)
        v = vocab
        self.source_type = CodeCol(v.source_type.values, v.source_type.none_code, np.uint8)
        self.document_payer = CodeCol(v.payer.values, v.payer.none_code, np.uint16)
        self.document_state = CodeCol(v.state.values, v.state.none_code, np.uint8)
        self.document_program = CodeCol(v.program.values, v.program.none_code, np.uint8)
        self.document_authority_level = CodeCol(v.authority.values, v.authority.none_code, np.uint8)
        self.extra: dict[str, Any] = {c: (IntCol() if c in _INT_COLS else JsonCol() if c in _JSON_COLS else StrCol())
                                      for c in HYDRATE_COLS}
        self.doc_idx: dict[str, int] = {}            # live document_id -> doc index
        self.doc_ids: list[str] = []                 # doc index -> document_id (indices are never reused)
        self.doc_d_tags: dict[str, set] = {}
        self.doc_p_tags: dict[str, set] = {}
        self.doc_j_tags: dict[str, list] = {}
        self.dead_docs: set[int] = set()             # doc indices whose rows were deleted (a re-published document gets a new index)
        self.doc_rows = np.zeros(1024, dtype=np.int64)   # rows per doc index (count(*) .. WHERE document_id = ANY(..), corpus_search.py:3461-3470)
        self.doc_listeners: list = []                # callables(doc_idx): a document's tags changed (derived row features go stale)

    def __len__(self) -> int:
        return self._n

    # -- write side (publish.py:204-362 / embedding_worker.py:229-266) -------------------------
    def _doc(self, document_id: str) -> int:
        d = self.doc_idx.get(document_id)
        if d is None:
            d = len(self.doc_ids)
            self.doc_ids.append(document_id)
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
        valid = np.fromiter((e is not None for e in embeddings), dtype=np.uint8, count=n)
        try:
            if valid.all():
                X = to_float4(embeddings)
            else:
                X = np.zeros((n, dim), dtype=np.float32)
                for i, e in enumerate(embeddings):
                    if e is not None:
                        v = to_float4(e)
                        if v.ndim != 1 or v.shape[0] != dim:
                            raise ValueError(f"expected {dim} dimensions, not {v.shape[0] if v.ndim == 1 else v.shape}")
                        X[i] = v
        except ValueError as exc:
            if "inhomogeneous" in str(exc):
                raise ValueError(f"expected {dim} dimensions for every embedding") from None
            raise
        if X.ndim != 2 or X.shape[1] != dim:
            raise ValueError(f"expected {dim} dimensions, not {X.shape[1] if X.ndim == 2 else X.shape}")   # pgvector's error text
        cols = {name: [r.get(name) for r in rows] for name in
                ("id", "document_id", "source_type", "source_id", "document_payer", "document_state", "document_program",
                 "document_authority_level", *HYDRATE_COLS)}
        self.insert_columns(X, cols, valid)

    def insert_columns(self, X: np.ndarray, columns: dict[str, Sequence], valid: np.ndarray | None = None) -> int:
        """Bulk INSERT: X float32 [n, dim] (rows with valid[i] == 0 have no vector), ``columns`` one sequence per
        column name (``id`` and ``document_id`` required; missing columns are NULL).  Returns the first new row."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        n = X.shape[0]
        if n == 0:
            return self._n
        if X.ndim != 2 or X.shape[1] != self.index.dim:
            raise ValueError(f"expected {self.index.dim} dimensions, not {X.shape[1] if X.ndim == 2 else X.shape}")
        for name in ("id", "document_id"):
            if name not in columns or len(columns[name]) != n:
                raise ValueError(f"column {name!r} is required, one value per row")
        none = [None] * n
        with self.lock:
            v = self.vocab
            first = self._n
            docs = np.fromiter((self._doc(str(d)) for d in columns["document_id"]), dtype=np.uint32, count=n)
            codes = {}
            for col, vname, dt in _CODED:
                voc = getattr(v, vname)
                codes[col] = np.fromiter((voc.encode(x) for x in columns.get(col, none)), dtype=dt, count=n)
            meta = make_meta(n, doc_idx=docs, payer=codes["document_payer"], state=codes["document_state"],
                             program=codes["document_program"], authority=codes["document_authority_level"],
                             source_type=codes["source_type"], valid=np.ones(n, np.uint8) if valid is None else valid)
            # 1. host columns first ...
            self.id.extend(str(x) for x in columns["id"])
            self.source_id.extend(None if x is None else str(x) for x in columns.get("source_id", none))
            if first + n > self.row_doc.shape[0]:
                self.row_doc = np.concatenate([self.row_doc, np.zeros(max(first + n, self.row_doc.shape[0]), dtype=np.uint32)])
            self.row_doc[first:first + n] = docs
            self._host_n = first + n
            if len(self.doc_ids) > self.doc_rows.shape[0]:
                self.doc_rows = np.concatenate([self.doc_rows, np.zeros(max(len(self.doc_ids), self.doc_rows.shape[0]), dtype=np.int64)])
            np.add.at(self.doc_rows, docs, 1)
            for col, _, _ in _CODED:
                getattr(self, col).extend(codes[col])
            for c in HYDRATE_COLS:
                self.extra[c].extend(columns.get(c, none))
            # 2. ... then the device rows: from here on a search may return them
            try:
                got = self.index.append(X, meta)
            except Exception:
                self._truncate(first)
                raise
            assert got == first, "host table and device index out of step"
            self._n = first + n
            return first

    def _truncate(self, n: int) -> None:
        self._host_n = n
        for col in (self.id, self.source_id, self.source_type, self.document_payer, self.document_state,
                    self.document_program, self.document_authority_level, *self.extra.values()):
            col.truncate(n)

    def set_document_j_tags(self, document_id: str, j_tags: Sequence[str] | None) -> None:
        """document_tags.j_tags of one document (app/models.py:535-537): binary coverage credit of `_rerank` and the J side
        of the candidate-pool cascade."""
        with self.lock:
            self.doc_j_tags[str(document_id)] = list(j_tags or ())
            bits = np.zeros((1, N.MRAG_JTAG_WORDS), dtype=np.uint64)
            for key in j_tags or ():
                b = self.vocab.jtag_bit(key, allocate=True)
                bits[0, b >> 6] |= np.uint64(1 << (b & 63))
            d = self._doc(str(document_id))
            self.index.set_doc_jtags(d, bits)
            for fn in self.doc_listeners:
                fn(d)

    def set_document_tags(self, document_id: str, d_tags: Sequence[str] | None, p_tags: Sequence[str] | None,
                          j_tags: Sequence[str] | None = None) -> None:
        """UPSERT one document_tags row (keys of d_tags / p_tags / j_tags, app/models.py:525-543)."""
        if j_tags is not None:
            self.set_document_j_tags(document_id, j_tags)
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
            for fn in self.doc_listeners:
                fn(d)

    def delete_document(self, document_id: str) -> int:
        """DELETE FROM .. WHERE document_id = :id (publish.py:310-313)."""
        with self.lock:
            d = self.doc_idx.get(str(document_id))
            if d is None:
                return 0
            n = self.index.tombstone_doc(d)
            # a re-published document gets a fresh doc_idx so the tombstoned rows stay dead; its document_tags row goes
            # with it (host dictionaries and the device tag sets the pool cascade reads)
            del self.doc_idx[str(document_id)]
            self.dead_docs.add(d)
            if self.doc_d_tags.pop(str(document_id), None) is not None or self.doc_p_tags.pop(str(document_id), None) is not None:
                self.index.set_doc_tags(d, np.zeros((1, N.MRAG_TAG_WORDS), dtype=np.uint64))
            self.doc_p_tags.pop(str(document_id), None)
            if self.doc_j_tags.pop(str(document_id), None) is not None:
                self.index.set_doc_jtags(d, np.zeros((1, N.MRAG_JTAG_WORDS), dtype=np.uint64))
            return n

    # -- snapshot ------------------------------------------------------------------------------
    def save(self, dirpath: str, corpus_version: int = 0) -> None:
        """Snapshot = the device shard(s) (``index.mrag*``) + the host columns (``columns.npz``, plain arrays) + the small
        dictionaries (``table.json``).  Keyed by corpus_state.corpus_version (publish.py:314).  No pickle anywhere."""
        os.makedirs(dirpath, exist_ok=True)
        with self.lock:
            self.index.save(os.path.join(dirpath, "index.mrag"), corpus_version)
            arrays = {"row_doc": self.row_doc[:self._n].copy()}
            arrays.update(self.id.arrays("id"))
            arrays.update(self.source_id.arrays("source_id"))
            for col, _, _ in _CODED:
                arrays[col + ".codes"] = getattr(self, col).view().copy()
            for c in HYDRATE_COLS:
                arrays.update(self.extra[c].arrays("extra." + c))
            np.savez(os.path.join(dirpath, "columns.npz"), **arrays)
            v = self.vocab
            meta = {
                "corpus_version": int(corpus_version), "n": self._n, "dim": self.index.dim,
                "shards": len(getattr(self.index, "shards", [None])),
                "vocab": {name: getattr(v, name).values for name in ("payer", "state", "program", "authority", "source_type")},
                "tag_bits": [[k[0], k[1], b] for k, b in v._tag_bit.items()], "jtag_bits": v._jtag_bit,
                "doc_j_tags": self.doc_j_tags,
                "doc_ids": self.doc_ids, "live_docs": sorted(self.doc_idx.values()), "dead_docs": sorted(self.dead_docs),
                "doc_d_tags": {k: sorted(s) for k, s in self.doc_d_tags.items()},
                "doc_p_tags": {k: sorted(s) for k, s in self.doc_p_tags.items()},
            }
            with open(os.path.join(dirpath, "table.json"), "w") as f:
                json.dump(meta, f)

    @classmethod
    def load(cls, dirpath: str, device: int = 0, capacity: int = 0, devices: Sequence[int] | None = None) -> tuple["PublishedTable", int]:
        """(table, corpus_version) from a snapshot directory; the cold start is one sequential read."""
        with open(os.path.join(dirpath, "table.json")) as f:
            meta = json.load(f)
        self = cls.__new__(cls)
        path = os.path.join(dirpath, "index.mrag")
        if meta["shards"] > 1 or (devices is not None and len(devices) > 1):
            from .multi import MultiIndex
            devs = list(devices) if devices is not None else [device] * meta["shards"]
            if len(devs) != meta["shards"]:
                raise ValueError(f"snapshot holds {meta['shards']} shards, {len(devs)} devices given")
            self.index, ver = MultiIndex.load(path, devs, capacity)
        else:
            self.index, ver = Index.load(path, device if devices is None else devices[0], capacity)
        if ver != meta["corpus_version"] or len(self.index) != meta["n"]:
            self.index.close()
            raise ValueError("snapshot is inconsistent: index.mrag and table.json come from different versions")
        v = Vocab()
        for name, values in meta["vocab"].items():
            voc = getattr(v, name)
            for x in values:
                voc.encode(x)
        for kind, key, b in meta["tag_bits"]:
            v._tag_bit[(kind, key)] = int(b)
        v._jtag_bit = {k: int(b) for k, b in meta.get("jtag_bits", {}).items()}
        self._init_host(v)
        z = np.load(os.path.join(dirpath, "columns.npz"), allow_pickle=False)
        self._n = self._host_n = int(meta["n"])
        self.row_doc = np.ascontiguousarray(z["row_doc"], dtype=np.uint32)
        self.id = StrCol.from_arrays(z, "id")
        self.source_id = StrCol.from_arrays(z, "source_id")
        for col, _, _ in _CODED:
            getattr(self, col).extend(z[col + ".codes"])
        for c in HYDRATE_COLS:
            kind = IntCol if c in _INT_COLS else JsonCol if c in _JSON_COLS else StrCol
            self.extra[c] = kind.from_arrays(z, "extra." + c)
        self.doc_ids = list(meta["doc_ids"])
        self.doc_idx = {self.doc_ids[d]: d for d in meta["live_docs"]}
        self.dead_docs = set(int(d) for d in meta.get("dead_docs", []))
        self.doc_rows = np.bincount(self.row_doc[:self._n], minlength=max(len(self.doc_ids), 1024)).astype(np.int64)
        self.doc_d_tags = {k: set(s) for k, s in meta["doc_d_tags"].items()}
        self.doc_p_tags = {k: set(s) for k, s in meta["doc_p_tags"].items()}
        self.doc_j_tags = {k: list(s) for k, s in meta.get("doc_j_tags", {}).items()}
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
        if include_document_ids is not None and hasattr(include_document_ids, "doc_indices"):
            # a CandidatePool built on the device (pool.build_candidate_pool): its document bitmap is used as it is
            pool = include_document_ids
            if pool._handle is not None and not hasattr(self.index, "shards"):
                f.pool_handle(pool._handle)
            else:
                f.doc_pool(pool.doc_indices())
        elif include_document_ids:
            f.doc_pool([d for d in (self.doc_idx.get(str(x)) for x in include_document_ids) if d is not None])
        return f
