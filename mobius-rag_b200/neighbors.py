"""Sibling-paragraph index: `_fetch_sibling_chunks_batch` without SQL.

The reference fetches the +-N paragraphs (within +-M pages) around each seed chunk with one UNNEST join over
``rag_published_embeddings`` (app/services/corpus_search.py:2560-2687) and uses the result three times per search: to
enrich the rerank haystacks (:2823-2918), to merge a topic block into a high-similarity seed (:2921-3076) and to expand
the assembled answer with context (:3079-3184).  Here the table's rows are already in host columns, so the join is a
lookup in a (document, paragraph_index) sorted index:

    build   one lexsort over (doc_idx, paragraph_index) of the live rows that have both a paragraph and a page number
    fetch   per seed two binary searches (document range, paragraph window) + a page-window filter on a handful of rows

The statement's semantics are kept: a seed is excluded only from ITS OWN window (another seed's window may return it),
rows with a NULL paragraph / page never match BETWEEN, DISTINCT ON (id) ORDER BY id, LIMIT 500.
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from .table import PublishedTable

NO_PAGE_HI = 10_000_000          # a seed without an integer page gets no page constraint (:2557)
FETCH_LIMIT = 500


def _none_if_empty(v):
    if v is None:
        return None
    s = str(v).strip()
    return s or None


class NeighborIndex:
    def __init__(self, table: PublishedTable):
        self.table = table
        self._stamp = None
        self._order = self._doc = self._para = self._page = None

    def _ensure(self) -> None:
        t = self.table
        stamp = (len(t), len(t.dead_docs))
        if stamp == self._stamp:
            return
        n = stamp[0]
        para, page = t.extra["paragraph_index"], t.extra["page_number"]
        ok = ~para.null[:n] & ~page.null[:n]
        if t.dead_docs:
            ok &= ~np.isin(t.row_doc[:n], np.fromiter(t.dead_docs, dtype=np.uint32))
        rows = np.flatnonzero(ok)
        d, p = t.row_doc[rows].astype(np.int64), para.val[rows]
        order = np.lexsort((p, d))
        self._order, self._doc, self._para = rows[order], d[order], p[order]
        self._page = page.val[rows][order]
        self._stamp = stamp

    def rows_near(self, doc_idx: int, para_lo: int, para_hi: int, page_lo: int, page_hi: int) -> np.ndarray:
        """table rows of document `doc_idx` with paragraph_index in [para_lo, para_hi] and page_number in [page_lo, page_hi]"""
        self._ensure()
        a, b = np.searchsorted(self._doc, doc_idx, "left"), np.searchsorted(self._doc, doc_idx, "right")
        if a == b:
            return np.zeros(0, dtype=np.int64)
        seg = self._para[a:b]
        i0, i1 = a + np.searchsorted(seg, para_lo, "left"), a + np.searchsorted(seg, para_hi, "right")
        pg = self._page[i0:i1]
        return self._order[i0:i1][(pg >= page_lo) & (pg <= page_hi)]

    def fetch_siblings(self, seeds: Sequence[dict[str, Any]], *, paragraph_window: int = 2, page_window: int = 1) -> list[dict[str, Any]]:
        """`_fetch_sibling_chunks_batch`: chunk-shaped dicts (``is_neighbor=True``) around every seed, seeds' own ids excluded
        from their own windows."""
        t = self.table
        found: set[int] = set()
        for s in seeds:
            doc_id = s.get("document_id")
            if not doc_id:
                continue
            d = t.doc_idx.get(str(doc_id))
            if d is None:
                continue
            pi = s.get("paragraph_index")
            pi = int(pi) if pi is not None else 0
            page = s.get("page_number")
            if isinstance(page, int):
                plo, phi = max(0, page - page_window), page + page_window
            else:
                plo, phi = 0, NO_PAGE_HI
            own = str(s.get("id")) if s.get("id") is not None else ""
            for r in self.rows_near(d, max(0, pi - paragraph_window), pi + paragraph_window, plo, phi).tolist():
                if r in found or (own and t.id[r] == own):
                    continue
                found.add(r)
        picked = sorted(found, key=lambda r: t.id[r])[:FETCH_LIMIT]          # DISTINCT ON (id) ORDER BY id LIMIT 500
        ex = t.extra
        out = []
        for r in picked:
            out.append({
                "id": t.id[r], "document_id": t.document_id[r], "text": ex["text"][r] or "",
                "page_number": ex["page_number"][r], "paragraph_index": ex["paragraph_index"][r],
                "section_path": _none_if_empty(ex["section_path"][r]), "chapter_path": _none_if_empty(ex["chapter_path"][r]),
                "summary": _none_if_empty(ex["summary"][r]), "content_sha": _none_if_empty(ex["content_sha"][r]),
                "document_name": ex["document_display_name"][r] or ex["document_filename"][r] or "document",
                "source_type": "hierarchical", "similarity": 0.0, "rerank_score": 0.0, "confidence_label": "low",
                "retrieval_arms": ["neighbor"], "authority_level": _none_if_empty(t.document_authority_level[r]),
                "payer": _none_if_empty(t.document_payer[r]), "state": _none_if_empty(t.document_state[r]),
                "jpd_tags": [], "is_neighbor": True,
            })
        return out
