"""The vector arm of corpus_search, mirroring ``_vector_arm`` (app/services/corpus_search.py:1427-1602).

Same signature (the first argument is the PublishedTable instead of an AsyncSession), same
LIMIT arithmetic, strict -> relaxed tag-filter retry, clamp / min_similarity / stop-at-k
post-processing, same result dict keys (``_row_to_base_dict`` :563-587 + similarity,
match_score, _arm) and the same fail-soft rule: any exception is logged and ``[]`` returned.
"""
from __future__ import annotations

import asyncio
import os
import json
import logging
from dataclasses import dataclass, field
from typing import Any

from . import _native as N
from .index import Filter
from .table import PublishedTable, to_float4

logger = logging.getLogger(__name__)


@dataclass
class CorpusFilters:
    """corpus_search.py:74-78."""
    payer: str | None = None
    state: str | None = None
    program: str | None = None
    authority_level: str | None = None

    def __bool__(self) -> bool:   # pydantic models are always truthy; keep that
        return True


@dataclass
class LexiconExpansion:
    """corpus_search_lexicon.py:76-101 (the fields the arm reads)."""
    matched_codes: list[str] = field(default_factory=list)
    expansion_phrases: list[str] = field(default_factory=list)
    domain_tags: list[str] = field(default_factory=list)
    jurisdiction_tags: list[str] = field(default_factory=list)
    process_tags: list[str] = field(default_factory=list)
    log: list[str] = field(default_factory=list)


def _log_stage(stage: str, search_id: str, **fields: Any) -> None:
    """corpus_search.py:495-509: one structured line per stage."""
    if not search_id:
        return
    try:
        logger.info("corpus_search %s", json.dumps({"stage": stage, "search_id": search_id, **fields}, default=str))
    except Exception:   # logging never breaks a search
        pass


def _none_if_empty(v):
    if v is None:
        return None
    s = str(v).strip()
    return s or None


def _row_to_base_dict(t: PublishedTable, r: int) -> dict[str, Any]:
    """corpus_search.py:563-587."""
    ex = t.extra
    doc_name = (ex["document_display_name"][r] or "").strip() or ex["document_filename"][r] or ""
    return {
        "id": str(t.id[r]),
        "text": ex["text"][r] or "",
        "document_id": str(t.document_id[r]),
        "document_name": doc_name,
        "page_number": ex["page_number"][r],
        "paragraph_index": ex["paragraph_index"][r],
        "source_type": t.source_type[r] or "hierarchical",
        "authority_level": (t.document_authority_level[r] or "").strip() or None,
        "payer": (t.document_payer[r] or "").strip() or None,
        "state": (t.document_state[r] or "").strip() or None,
        "section_path": _none_if_empty(ex["section_path"][r]),
        "chapter_path": _none_if_empty(ex["chapter_path"][r]),
        "summary": _none_if_empty(ex["summary"][r]),
        "content_sha": _none_if_empty(ex["content_sha"][r]),
        "chunk_d_tags": ex["chunk_d_tags"][r] or {},
        "chunk_p_tags": ex["chunk_p_tags"][r] or {},
        "chunk_j_tags": ex["chunk_j_tags"][r] or {},
    }


def _tag_filters(t: PublishedTable, expansion: LexiconExpansion | None, tag_mode: str):
    """strict / relaxed tag clauses (corpus_search.py:1464-1510) as code sets.
    Returns (strict, relaxed); each is None when that SQL fragment would be empty, else a
    callable adding the clause to a Filter."""
    if expansion is None or (tag_mode or "auto").lower() == "none":
        return None, None

    def strip(tag: str, prefix: str):
        return tag[len(prefix):] if tag.startswith(prefix) else None

    j_keys = [k for k in (strip(x, "j:") for x in expansion.jurisdiction_tags) if k]
    d_keys = [k for k in (strip(x, "d:") for x in expansion.domain_tags) if k]
    p_keys = [k for k in (strip(x, "p:") for x in expansion.process_tags) if k]
    v = t.vocab
    states: list[int] = []
    programs: list[int] = []
    payers: list[int] = []
    n_clauses = 0
    for jk in j_keys:
        if "." not in jk:
            continue
        cat, val = jk.split(".", 1)
        val_human = val.replace("_", " ")
        if cat == "state":
            states.append(v.state.lookup(val.upper()[:2]))
            n_clauses += 1
        elif cat == "program":
            programs.extend(v.program.ilike(f"%{val_human}%"))
            n_clauses += 1
        elif cat in ("payor", "regulatory_authority"):
            payers.extend(v.payer.ilike(f"%{val_human}%"))
            n_clauses += 1
    strict = relaxed = None
    if n_clauses:
        strict = lambda f: f.tag_strict(states, programs, payers)
    if d_keys or p_keys:
        bits = [b for b in ([v.tag_bit("d", k, False) for k in d_keys] + [v.tag_bit("p", k, False) for k in p_keys])
                if b is not None]
        relaxed = lambda f: f.tag_relaxed(bits)
    return strict, relaxed


def vector_arm(
    table: PublishedTable,
    query_embedding: list[float],
    k: int,
    filters: CorpusFilters | None,
    include_document_ids: list[str] | None,
    search_id: str = "",
    expansion: LexiconExpansion | None = None,
    tag_mode: str = "auto",
    min_similarity: float | None = None,
    over_fetch_factor: int = 1,
) -> list[dict[str, Any]]:
    sql_limit = max(1, k) * max(1, int(over_fetch_factor))
    try:
        if sql_limit > N.MRAG_MAX_K:
            raise ValueError(f"LIMIT {sql_limit} exceeds MRAG_MAX_K={N.MRAG_MAX_K}")
        q = to_float4(query_embedding)
        if q.shape[0] != table.index.dim:
            raise ValueError(f"different vector dimensions {table.index.dim} and {q.shape[0]}")
        strict, relaxed = _tag_filters(table, expansion, tag_mode)
        tm = (tag_mode or "auto").lower().strip()
        if tm == "none":
            first, retry = None, None
        elif tm == "relaxed":
            first, retry = relaxed, None
        elif tm == "strict":
            first, retry = strict, None
        else:  # auto
            first, retry = strict, relaxed

        def run(tagclause):
            f: Filter = table.filter_corpus(filters, include_document_ids)
            if tagclause is not None:
                tagclause(f)
            opts = N.OPT_COALESCE if (not f.active and os.getenv("MRAG_COALESCE", "1") != "0") else 0
            scores, rows, counts = table.index.search(q[None, :], sql_limit, f if f.active else None, options=opts)
            n = int(counts[0])
            return rows[0, :n], scores[0, :n]

        rows, sims = run(first)
        # "if not rows_check and tag_filter_relaxed and tag_filter_relaxed != tag_filter_sql"
        if len(rows) == 0 and retry is not None:
            _log_stage("vector_tag_filter_relaxed", search_id,
                       note="strict returned 0; retrying with relaxed (d/p only)")
            rows, sims = run(retry)
    except Exception as exc:
        logger.error("corpus_search vector arm failed: %s", exc, exc_info=True)
        return []

    out: list[dict[str, Any]] = []
    n_below_threshold = 0
    for rank0, (r, s) in enumerate(zip(rows, sims)):
        cosine_sim = max(0.0, min(1.0, float(float(s) or 0.0)))          # corpus_search.py:1569
        if min_similarity is not None and cosine_sim < min_similarity:
            n_below_threshold += 1
            continue
        c = _row_to_base_dict(table, int(r))
        c["similarity"] = cosine_sim
        c["match_score"] = cosine_sim
        c["_arm"] = "vector"
        out.append(c)
        if len(out) >= k:
            break
        _log_stage("vector_arm", search_id, rank=rank0 + 1, chunk_id=c["id"], doc=c["document_name"][:50],
                   page=c["page_number"], authority=c["authority_level"], cosine=cosine_sim,
                   preview=(c["text"] or "")[:80])
    if search_id:
        _log_stage("vector_arm_summary", search_id, hits=len(out), scanned=len(rows),
                   below_threshold=n_below_threshold, threshold=min_similarity, sql_limit=sql_limit)
    return out


async def _vector_arm(db: PublishedTable, query_embedding, k, filters, include_document_ids, search_id: str = "",
                      expansion=None, tag_mode: str = "auto", min_similarity=None, over_fetch_factor: int = 1):
    """Coroutine with the reference's exact name and argument order, for callers that ``await`` it
    from asyncio tasks (corpus_search.py:3393-3401, 3506-3514)."""
    return await asyncio.to_thread(vector_arm, db, query_embedding, k, filters, include_document_ids, search_id,
                                   expansion, tag_mode, min_similarity, over_fetch_factor)
