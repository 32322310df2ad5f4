"""Vector-store plugin boundary, mirroring app/services/vector_store.py of the reference.

``B200VectorStore`` is the drop-in for ``PgVectorStore`` (vector_store.py:120-303): same method
names, arguments, result dict keys and error behaviour; the statement it used to send to
pgvector runs as CUDA kernels on a B200 instead.  ``get_vector_store()`` recognises
``VECTOR_STORE=b200`` next to the reference's own values (vector_store.py:306-321).
"""
from __future__ import annotations

import asyncio
import logging
import os
import threading
from abc import ABC, abstractmethod

from .table import PublishedTable

logger = logging.getLogger(__name__)


class VectorStore(ABC):
    """Abstract vector store for embeddings (vector_store.py:15-31)."""

    @abstractmethod
    def add(self, ids: list[str], embeddings: list[list[float]], metadata: list[dict]) -> None:
        """Add embeddings with ids and metadata. metadata items: document_id, source_type, source_id."""

    @abstractmethod
    def search(self, embedding: list[float], k: int = 10, document_id: str | None = None) -> list[dict]:
        """Return top-k results. Each result: {id, document_id, source_type, source_id, distance}."""

    @abstractmethod
    def delete_by_document(self, document_id: str) -> None:
        """Delete all embeddings for a document."""


class NoopVectorStore(VectorStore):
    """vector_store.py:107-117."""

    def add(self, ids, embeddings, metadata) -> None:
        pass

    def search(self, embedding, k: int = 10, document_id: str | None = None) -> list[dict]:
        return []

    def delete_by_document(self, document_id: str) -> None:
        pass


class B200VectorStore(VectorStore):
    """Exact cosine top-k over a GPU-resident copy of the table.

    Unlike ``PgVectorStore`` (whose add/delete are no-ops because Postgres is the store), this
    store owns its rows, so ``add`` / ``delete_by_document`` do the work Chroma's did
    (vector_store.py:62-75, 101-104).  ``metadata`` items may carry, besides document_id /
    source_type / source_id, the denormalised filter columns of rag_published_embeddings
    (document_payer, document_state, document_authority_level, document_program) and the
    hydration columns; missing ones are NULL.
    """

    def __init__(self, table_name: str = "rag_published_embeddings", dim: int | None = None,
                 dtype: str | None = None, device: int | None = None, capacity: int | None = None,
                 table: PublishedTable | None = None, devices: list[int] | None = None):
        if table_name not in {"rag_published_embeddings", "chunk_embeddings"}:
            raise ValueError(f"B200VectorStore: unsupported table {table_name!r}")   # vector_store.py:163-164
        self._table_name = table_name
        self._dim = int(dim if dim is not None else os.getenv("EMBEDDING_DIMENSIONS", "1536"))
        self._dtype = (dtype or os.getenv("MRAG_DTYPE", "f32")).lower()
        self._device = int(device if device is not None else os.getenv("MRAG_DEVICE", "0"))
        self._capacity = int(capacity if capacity is not None else os.getenv("MRAG_CAPACITY", str(1 << 21)))
        # MRAG_DEVICES="0,1,2,3,4,5,6,7": ONE store object (one FastAPI worker) row-shards the table over these GPUs;
        # search / asearch keep the reference's signatures (vector_store.py:181-226) and hit all of them
        env_devs = os.getenv("MRAG_DEVICES", "").strip()
        self._devices = list(devices) if devices else ([int(x) for x in env_devs.split(",") if x.strip()] if env_devs else None)
        self._table = table
        self._init_lock = threading.Lock()

    @property
    def table(self) -> PublishedTable:
        if self._table is None:
            with self._init_lock:
                if self._table is None:
                    self._table = PublishedTable(self._dim, self._dtype, self._device, self._capacity, devices=self._devices)
        return self._table

    def add(self, ids: list[str], embeddings: list[list[float]], metadata: list[dict]) -> None:
        if not ids:
            return
        rows = []
        for id_, m in zip(ids, metadata):
            r = dict(m)
            r["id"] = id_
            r["document_id"] = str(m.get("document_id", ""))
            r["source_type"] = m.get("source_type")
            r["source_id"] = m.get("source_id")
            rows.append(r)
        self.table.insert(rows, embeddings)
        logger.debug("B200VectorStore: added %d embeddings", len(ids))

    def delete_by_document(self, document_id: str) -> None:
        n = self.table.delete_document(document_id)
        logger.debug("B200VectorStore: deleted %d embeddings for document %s", n, document_id)

    def search(self, embedding: list[float], k: int = 10, document_id: str | None = None,
               filters: dict[str, str] | None = None) -> list[dict]:
        """Same contract as PgVectorStore.search (vector_store.py:181-216): refuses to block a
        running event loop; ``distance`` holds cosine SIMILARITY (1.0 = identical)."""
        try:
            asyncio.get_running_loop()
        except RuntimeError:
            return self._search_blocking(embedding, k, document_id, filters)
        raise RuntimeError(
            "B200VectorStore.search() called from a running event loop; "
            "use ``await store.asearch(...)`` instead."
        )

    async def asearch(self, embedding: list[float], k: int = 10, document_id: str | None = None,
                      filters: dict[str, str] | None = None) -> list[dict]:
        """Async variant (vector_store.py:218-226).  The scan releases the GIL inside the C ABI."""
        return await asyncio.to_thread(self._search_blocking, embedding, k, document_id, filters)

    def _search_blocking(self, embedding, k, document_id, filters) -> list[dict]:
        from .table import to_float4
        t = self.table
        k = int(k)
        if k < 1:
            raise ValueError("LIMIT must not be negative" if k < 0 else "k must be >= 1")
        q = to_float4(embedding)
        if q.shape[0] != t.index.dim:
            raise ValueError(f"different vector dimensions {t.index.dim} and {q.shape[0]}")   # pgvector's error
        flt = t.filter_pg_store(document_id, filters)
        # unfiltered requests from concurrent threads (asearch -> to_thread) share one pass over the corpus
        from . import _native as N
        opts = N.OPT_COALESCE if (flt is None and os.getenv("MRAG_COALESCE", "1") != "0") else 0
        scores, rows, counts = t.index.search(q[None, :], k, flt, options=opts)
        out: list[dict] = []
        for j in range(int(counts[0])):
            r = int(rows[0, j])
            out.append({
                "id": t.id[r],
                "document_id": t.document_id[r],
                "source_type": t.source_type[r],
                "source_id": t.source_id[r],
                # Match Chroma's key; value is similarity per spec (vector_store.py:300-301).
                "distance": float(scores[0, j]),
            })
        return out


class B200ChromaVectorStore(B200VectorStore):
    """Drop-in for ``ChromaVectorStore`` (vector_store.py:34-104): the collection ``coll.query(query_embeddings=[e],
    n_results=k, where={"document_id": ...}, include=[metadatas, distances])`` on a ``hnsw:space = cosine`` collection.

    Differences from the pgvector-shaped store, all the reference's: ``distance`` is the cosine DISTANCE (0..2, smaller is
    closer), not the similarity; the only filter is ``document_id``; metadata values are stringified on ``add``
    (``str(m.get(key, ""))``, :70-74 -- a None becomes the string "None").  Chroma's HNSW search is approximate; this scan
    is exact, i.e. it returns what Chroma would with perfect recall.  A zero-norm vector has no cosine; hnswlib reports
    distance 1.0 for it and so does this store."""

    def __init__(self, collection_name: str = "chunk_embeddings", host: str | None = None, port: int | None = None,
                 persist_directory: str | None = None, **kw):
        super().__init__(table_name="chunk_embeddings", **kw)
        self._collection_name = collection_name
        self._host, self._port, self._persist_directory = host, port, persist_directory     # kept for signature parity; unused

    def add(self, ids: list[str], embeddings: list[list[float]], metadata: list[dict]) -> None:
        if not ids:
            return
        metas = [{"document_id": str(m.get("document_id", "")), "source_type": str(m.get("source_type", "")),
                  "source_id": str(m.get("source_id", ""))} for m in metadata]
        super().add(ids, embeddings, metas)

    def search(self, embedding: list[float], k: int = 10, document_id: str | None = None) -> list[dict]:
        out = self._search_blocking(embedding, k, document_id, None)
        for r in out:
            sim = r["distance"]
            r["distance"] = 1.0 if sim != sim else 1.0 - sim
        return out

    async def asearch(self, embedding: list[float], k: int = 10, document_id: str | None = None, filters=None) -> list[dict]:
        return await asyncio.to_thread(self.search, embedding, k, document_id)


def get_vector_store() -> VectorStore:
    """vector_store.py:306-321 with the B200 stores in place of the engines: ``VECTOR_STORE=b200`` -> the pgvector-shaped
    B200VectorStore; Chroma's environment (``CHROMA_HOST`` / ``CHROMA_PERSIST_DIR``) -> the Chroma-shaped store; otherwise
    the no-op store, as in the reference.  ``VECTOR_STORE=pgvector`` is the reference's own Postgres path (INTEGRATION.md
    shows the one-line edit that adds the ``b200`` branch next to it)."""
    explicit = (os.getenv("VECTOR_STORE") or "").strip().lower()
    if explicit == "b200":
        return B200VectorStore()
    if explicit == "pgvector":
        raise RuntimeError("VECTOR_STORE=pgvector is the reference's Postgres path; this package provides VECTOR_STORE=b200")
    if os.getenv("CHROMA_HOST") or os.getenv("CHROMA_PERSIST_DIR"):
        return B200ChromaVectorStore()
    return NoopVectorStore()
