"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the retrieval hot path.

PARITY UNPINNED.  The reference (ananthlk/Mobius-RAG) delegates the arithmetic of this path to
pgvector (`<=>`), which is not vendored under /root/reference, cannot be installed here (no
Postgres, no network) and whose results the reference's own tests never pin.
This module therefore restates

  * pgvector v0.5.1 ``cosine_distance`` (published algorithm; the only pin in the reference is
    scripts/install_pgvector_for_postgresql14.sh:21) -- in C (oracle/pgv_oracle.c) and in numpy,
  * the SQL the reference wraps around it, clause by clause, evaluated on the *string* columns
    of ``rag_published_embeddings`` exactly as Postgres would (the product evaluates the same
    clauses on dictionary codes and bitsets, so the two share no code):
      - app/services/vector_store.py:245-303      PgVectorStore._search_async
      - app/services/corpus_search.py:516-560     _build_filter_clauses
      - app/services/corpus_search.py:1458-1579   _vector_arm (LIMIT, tag filters, retry, post-filter)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg may
import this package.  Nothing under mobius-rag_b200/ does.
"""
from __future__ import annotations

import ctypes
import math
import os
import re
import subprocess
from dataclasses import dataclass, field
from typing import Any, Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpgv_oracle.so")
_lib = None


def build_clib(force: bool = False) -> str:
    """Compile oracle/pgv_oracle.c (gcc, flags in oracle/Makefile)."""
    src = os.path.join(_HERE, "pgv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "libpgv_oracle.so"], check=True)
    return _LIB_PATH


def clib():
    """ctypes handle to the C restatement (built on first use; rebuilt if it will not load)."""
    global _lib
    if _lib is not None:
        return _lib
    build_clib()
    try:
        lib = ctypes.CDLL(_LIB_PATH)
    except OSError:
        build_clib(force=True)
        lib = ctypes.CDLL(_LIB_PATH)
    c_f32p = ctypes.POINTER(ctypes.c_float)
    c_f64p = ctypes.POINTER(ctypes.c_double)
    c_i64p = ctypes.POINTER(ctypes.c_int64)
    c_u8p = ctypes.POINTER(ctypes.c_uint8)
    lib.pgv_cosine_distance.restype = ctypes.c_double
    lib.pgv_cosine_distance.argtypes = [c_f32p, c_f32p, ctypes.c_int]
    lib.pgv_scan.restype = None
    lib.pgv_scan.argtypes = [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, c_u8p, c_f64p]
    lib.pgv_topk.restype = ctypes.c_int64
    lib.pgv_topk.argtypes = [c_f64p, c_u8p, ctypes.c_int64, ctypes.c_int64, c_i64p, c_f64p]
    lib.pgv_search_batch.restype = None
    lib.pgv_search_batch.argtypes = [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, ctypes.c_int,
                                     c_u8p, ctypes.c_int64, c_i64p, c_f64p, c_i64p, c_f64p]
    lib.pgv_set_threads.restype = None
    lib.pgv_set_threads.argtypes = [ctypes.c_int]
    lib.pgv_round_bf16.restype = None
    lib.pgv_round_bf16.argtypes = [c_f32p, c_f32p, ctypes.c_int64]
    _lib = lib
    return lib


def set_threads(t: int) -> None:
    """Number of scan threads of the C oracle (1 = one Postgres backend)."""
    clib().pgv_set_threads(int(t))


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ---------------------------------------------------------------------------------------------
# numeric format of the stored corpus
# ---------------------------------------------------------------------------------------------

def to_float4(emb: Iterable[float]) -> np.ndarray:
    """What pgvector stores for one embedding: the worker sends ``repr(float(x))`` text
    (app/embedding_worker.py:53-62, app/services/publish.py:337-341) and pgvector parses each
    element with strtof, i.e. every element is exactly ``np.float32(python_float)``."""
    return np.asarray([float(x) for x in emb], dtype=np.float64).astype(np.float32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32: what MRAG_BF16 storage keeps."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).reshape(x.shape)


# ---------------------------------------------------------------------------------------------
# pgvector `<=>`
# ---------------------------------------------------------------------------------------------

def cosine_distance_np(X: np.ndarray, q: np.ndarray) -> np.ndarray:
    """numpy restatement of pgvector cosine_distance for every row of X (float8 result).

    float32 accumulation of dot / |a|^2 / |b|^2 (summation order unspecified, as in the
    -fassociative-math build), float64 for the divide, clamp to [-1, 1], ``1 - sim``."""
    X = np.asarray(X, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    dot = (X @ q).astype(np.float32)
    na = np.einsum("ij,ij->i", X, X, dtype=np.float32)
    nb = np.float32(np.dot(q, q))
    with np.errstate(divide="ignore", invalid="ignore"):
        sim = dot.astype(np.float64) / np.sqrt(na.astype(np.float64) * np.float64(nb))
    sim = np.where(sim > 1, 1.0, np.where(sim < -1, -1.0, sim))
    return 1.0 - sim


def cosine_distance_c(X: np.ndarray, q: np.ndarray, mask: np.ndarray | None = None) -> np.ndarray:
    """Same through the C restatement (row loop exactly as src/vector.c)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, dim = X.shape
    dist = np.empty(n, dtype=np.float64)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    clib().pgv_scan(_p(X, ctypes.c_float), n, dim, dim, _p(q, ctypes.c_float),
                    None if m is None else _p(m, ctypes.c_uint8), _p(dist, ctypes.c_double))
    return dist


def order_by_limit(dist: np.ndarray, mask: np.ndarray | None, k: int) -> tuple[np.ndarray, np.ndarray]:
    """``ORDER BY dist ASC LIMIT k`` over rows where mask is true; NaN last (Postgres float8
    ordering); ties by ascending row.  Returns (rows int64[m], similarity float64[m]) with
    similarity = ``1 - dist`` evaluated in float8 like the SELECT list."""
    n = dist.shape[0]
    rows = np.arange(n, dtype=np.int64) if mask is None else np.nonzero(np.asarray(mask))[0].astype(np.int64)
    d = dist[rows]
    nan = np.isnan(d)
    order = np.lexsort((rows, np.where(nan, 0.0, d), nan))   # last key is primary
    order = order[: max(0, int(k))]
    rows = rows[order]
    return rows, 1.0 - d[order]


def search(X: np.ndarray, Q: np.ndarray, k: int, mask: np.ndarray | None = None, use_c: bool = True):
    """The statement for each query of Q.  Returns (rows[nq,k] int64 (-1 padded),
    sims[nq,k] float64 (NaN padded), counts[nq])."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(np.atleast_2d(Q), dtype=np.float32)
    n, dim = X.shape
    nq = Q.shape[0]
    rows = np.full((nq, k), -1, dtype=np.int64)
    sims = np.full((nq, k), np.nan, dtype=np.float64)
    counts = np.zeros(nq, dtype=np.int64)
    if use_c:
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        scratch = np.empty(max(n, 1), dtype=np.float64)
        clib().pgv_search_batch(_p(X, ctypes.c_float), n, dim, dim, _p(Q, ctypes.c_float), nq,
                                None if m is None else _p(m, ctypes.c_uint8), k,
                                _p(rows, ctypes.c_int64), _p(sims, ctypes.c_double),
                                _p(counts, ctypes.c_int64), _p(scratch, ctypes.c_double))
        return rows, sims, counts
    for i in range(nq):
        r, s = order_by_limit(cosine_distance_np(X, Q[i]), mask, k)
        rows[i, : len(r)] = r
        sims[i, : len(r)] = s
        counts[i] = len(r)
    return rows, sims, counts


def all_similarities(X: np.ndarray, q: np.ndarray) -> np.ndarray:
    """``1 - (v <=> q)`` for every row (float8), used by the tie-aware comparators."""
    return 1.0 - cosine_distance_c(X, q)


# ---------------------------------------------------------------------------------------------
# the table and the WHERE clauses (string level, as Postgres evaluates them)
# ---------------------------------------------------------------------------------------------

@dataclass
class Table:
    """Columns of rag_published_embeddings the path touches (app/models.py:242-280) plus the
    per-document tag aggregates of document_tags (app/models.py:525-543)."""
    id: list[str]
    document_id: list[str]
    source_type: list[str]
    source_id: list[str]
    document_payer: list[str]
    document_state: list[str]
    document_program: list[str]
    document_authority_level: list[str]
    has_vec: np.ndarray                       # embedding_vec IS NOT NULL
    X: np.ndarray                             # float4 vectors [n, dim] (rows without a vec are ignored)
    doc_d_tags: dict[str, set] = field(default_factory=dict)   # document_id -> keys of document_tags.d_tags
    doc_p_tags: dict[str, set] = field(default_factory=dict)   # document_id -> keys of document_tags.p_tags
    extra: dict[str, list] = field(default_factory=dict)       # text, page_number, ... for _row_to_base_dict

    def __len__(self) -> int:
        return len(self.id)

    def col(self, name: str) -> np.ndarray:
        return np.asarray(getattr(self, name), dtype=object)


def sql_ilike(value: str, pattern: str) -> bool:
    """Postgres ``value ILIKE pattern``: case-insensitive, ``%`` any run, ``_`` any one char."""
    if value is None:            # NULL ILIKE pattern is NULL, i.e. not true
        return False
    rx = "".join(".*" if c == "%" else "." if c == "_" else re.escape(c) for c in pattern)
    return re.fullmatch(rx, value, flags=re.IGNORECASE | re.DOTALL) is not None


# corpus_search.py:208-213
FL_MEDICAID_MCO_PAYERS = frozenset({
    "Sunshine Health", "Simply Healthcare", "United Healthcare",
    "Aetna", "Molina Healthcare", "Molina Healthcare of Florida",
    "WellCare", "Humana", "Humana Healthy Horizons",
})
FL_STATE_AUTHORITY_PAYERS = ["AHCA", "Ahca.myflorida", "Florida Medicaid"]

# vector_store.py:149-157
PG_ALLOWED_FILTERS = {
    "payer": "document_payer",
    "state": "document_state",
    "authority_level": "document_authority_level",
    "document_id": "document_id",
    "source_type": "source_type",
}


def where_pg_store(t: Table, document_id: str | None, filters: dict | None) -> np.ndarray:
    """WHERE of PgVectorStore._search_async (vector_store.py:245-267): document_id equality,
    whitelisted equality filters (None / '' / unknown keys skipped), embedding_vec IS NOT NULL."""
    m = np.asarray(t.has_vec, dtype=bool).copy()
    if document_id:
        m &= t.col("document_id") == document_id
    for key, value in (filters or {}).items():
        if value is None or value == "":
            continue
        colname = PG_ALLOWED_FILTERS.get(key)
        if not colname:
            continue
        m &= t.col(colname) == value
    return m


def where_filter_clauses(t: Table, filters: Any, include_document_ids: Sequence[str] | None) -> np.ndarray:
    """_build_filter_clauses (corpus_search.py:516-560).  `filters` has attributes payer, state,
    program, authority_level (CorpusFilters, corpus_search.py:74-78) or is None."""
    m = np.ones(len(t), dtype=bool)
    if filters:
        payer = getattr(filters, "payer", None)
        if payer:
            pm = t.col("document_payer") == payer
            if payer in FL_MEDICAID_MCO_PAYERS:
                pm |= np.isin(t.col("document_payer"), FL_STATE_AUTHORITY_PAYERS) & (t.col("document_state") == "FL")
            m &= pm
        if getattr(filters, "state", None):
            m &= t.col("document_state") == filters.state
        if getattr(filters, "program", None):
            m &= t.col("document_program") == filters.program
        if getattr(filters, "authority_level", None):
            m &= t.col("document_authority_level") == filters.authority_level
    if include_document_ids:
        m &= np.isin(t.col("document_id"), list(include_document_ids))
    return m


def tag_filter_masks(t: Table, expansion: Any, tag_mode: str) -> tuple[np.ndarray | None, np.ndarray | None]:
    """strict / relaxed tag filters of _vector_arm (corpus_search.py:1464-1510).
    Returns (strict_mask or None, relaxed_mask or None); None = that filter string is empty."""
    strict = relaxed = None
    if expansion is None or (tag_mode or "auto").lower() == "none":
        return None, None

    def strip(tag: str, prefix: str):
        return tag[len(prefix):] if tag.startswith(prefix) else None

    j_keys = [k for k in (strip(x, "j:") for x in expansion.jurisdiction_tags) if k]
    d_keys = [k for k in (strip(x, "d:") for x in expansion.domain_tags) if k]
    p_keys = [k for k in (strip(x, "p:") for x in expansion.process_tags) if k]
    clauses = []
    for jk in j_keys:
        if "." not in jk:
            continue
        cat, val = jk.split(".", 1)
        val_human = val.replace("_", " ")
        if cat == "state":
            clauses.append(t.col("document_state") == val.upper()[:2])
        elif cat == "program":
            pat = f"%{val_human}%"
            clauses.append(np.array([sql_ilike(s, pat) for s in t.document_program], dtype=bool))
        elif cat in ("payor", "regulatory_authority"):
            pat = f"%{val_human}%"
            clauses.append(np.array([sql_ilike(s, pat) for s in t.document_payer], dtype=bool))
    if clauses:
        strict = np.zeros(len(t), dtype=bool)
        for c in clauses:
            strict |= c
    if d_keys or p_keys:
        # LEFT JOIN document_tags dt ... jsonb_exists(dt.d_tags, key) OR jsonb_exists(dt.p_tags, key)
        # (a document without a document_tags row joins to NULLs -> jsonb_exists(NULL) is NULL -> not true)
        relaxed = np.zeros(len(t), dtype=bool)
        for i, did in enumerate(t.document_id):
            dt = t.doc_d_tags.get(did)
            pt = t.doc_p_tags.get(did)
            ok = False
            if dt is not None:
                ok = any(k in dt for k in d_keys)
            if not ok and pt is not None:
                ok = any(k in pt for k in p_keys)
            relaxed[i] = ok
    return strict, relaxed


# ---------------------------------------------------------------------------------------------
# the two statements
# ---------------------------------------------------------------------------------------------

def pg_store_search(t: Table, embedding: Sequence[float], k: int, document_id: str | None = None,
                    filters: dict | None = None) -> list[dict]:
    """PgVectorStore._search_async (vector_store.py:228-303) against Table `t`."""
    q = to_float4(embedding)
    mask = where_pg_store(t, document_id, filters)
    rows, sims = order_by_limit(cosine_distance_c(t.X, q), mask, k)
    out = []
    for r, s in zip(rows, sims):
        out.append({
            "id": t.id[r], "document_id": t.document_id[r], "source_type": t.source_type[r],
            "source_id": t.source_id[r], "distance": float(s),
        })
    return out


def none_if_empty(v):
    if v is None:
        return None
    s = str(v).strip()
    return s or None


def row_to_base_dict(t: Table, r: int) -> dict:
    """_row_to_base_dict (corpus_search.py:563-587)."""
    ex = t.extra

    def g(name, default=None):
        col = ex.get(name)
        return default if col is None else col[r]

    doc_name = (g("document_display_name", "") or "").strip() or g("document_filename", "") or ""
    return {
        "id": str(t.id[r]),
        "text": g("text", "") or "",
        "document_id": str(t.document_id[r]),
        "document_name": doc_name,
        "page_number": g("page_number"),
        "paragraph_index": g("paragraph_index"),
        "source_type": t.source_type[r] or "hierarchical",
        "authority_level": (t.document_authority_level[r] or "").strip() or None,
        "payer": (t.document_payer[r] or "").strip() or None,
        "state": (t.document_state[r] or "").strip() or None,
        "section_path": none_if_empty(g("section_path")),
        "chapter_path": none_if_empty(g("chapter_path")),
        "summary": none_if_empty(g("summary")),
        "content_sha": none_if_empty(g("content_sha")),
        "chunk_d_tags": g("chunk_d_tags") or {},
        "chunk_p_tags": g("chunk_p_tags") or {},
        "chunk_j_tags": g("chunk_j_tags") or {},
    }


def vector_arm(t: Table, query_embedding: Sequence[float], k: int, filters: Any = None,
               include_document_ids: Sequence[str] | None = None, expansion: Any = None,
               tag_mode: str = "auto", min_similarity: float | None = None,
               over_fetch_factor: int = 1) -> list[dict]:
    """_vector_arm (corpus_search.py:1427-1602) against Table `t`: LIMIT k*over_fetch, filter
    clauses, strict -> relaxed tag filter with retry on zero rows, clamp to [0,1], min_similarity
    post-filter, stop at k."""
    sql_limit = max(1, k) * max(1, int(over_fetch_factor))
    q = to_float4(query_embedding)
    base = np.asarray(t.has_vec, dtype=bool) & where_filter_clauses(t, filters, include_document_ids)
    strict, relaxed = tag_filter_masks(t, expansion, tag_mode)
    tm = (tag_mode or "auto").lower().strip()
    if tm == "none":
        first, retry = None, None
    elif tm == "relaxed":
        first, retry = relaxed, None
    elif tm == "strict":
        first, retry = strict, None
    else:
        first, retry = strict, relaxed
    dist = cosine_distance_c(t.X, q)

    def run(tagmask):
        m = base if tagmask is None else (base & tagmask)
        return order_by_limit(dist, m, sql_limit)

    rows, sims = run(first)
    # `tag_filter_relaxed != tag_filter_sql`: both are SQL strings; equal only if both empty
    if len(rows) == 0 and retry is not None:
        rows, sims = run(retry)
    out = []
    for r, s in zip(rows, sims):
        s = float(s)
        cosine_sim = max(0.0, min(1.0, float(s or 0.0)))      # corpus_search.py:1569 (NaN -> 1.0)
        if min_similarity is not None and cosine_sim < min_similarity:
            continue
        c = row_to_base_dict(t, int(r))
        c["similarity"] = cosine_sim
        c["match_score"] = cosine_sim
        c["_arm"] = "vector"
        out.append(c)
        if len(out) >= k:
            break
    return out


# ---------------------------------------------------------------------------------------------
# comparators (how "identical up to ties" is judged)
# ---------------------------------------------------------------------------------------------

def check_topk(rows_got: np.ndarray, scores_got: np.ndarray, count_got: int,
               sim_all: np.ndarray, mask: np.ndarray | None, k: int,
               rtol: float, tie_tol: float = 1e-6, atol_floor: float = 1e-3) -> None:
    """Assert that (rows_got, scores_got)[:count_got] is the statement's answer up to ties.

    sim_all = oracle ``1 - (v <=> q)`` for EVERY row (float8).  Checks
      1. count = min(k, rows passing the mask);
      2. no row repeated, every row passes the mask;
      3. position i holds a row whose oracle similarity equals the oracle's i-th best within
         tie_tol (so ids and order are identical except inside groups of near-equal scores);
         wherever the oracle's neighbours at i are separated by more than tie_tol the id
         itself must match;
      4. |score_got - oracle sim of that row| <= rtol * max(|sim|, atol_floor); NaN matches NaN.
    """
    n = sim_all.shape[0]
    m = np.ones(n, dtype=bool) if mask is None else np.asarray(mask, dtype=bool)
    ext_rows, ext_sims = order_by_limit(1.0 - sim_all, m, k + 1)   # one extra to see the boundary gap
    want_rows, want_sims = ext_rows[:k], ext_sims[:k]
    assert count_got == len(want_rows), f"count {count_got} != oracle {len(want_rows)}"
    got = np.asarray(rows_got[:count_got], dtype=np.int64)
    assert len(set(got.tolist())) == len(got), "duplicate rows in result"
    assert (got >= 0).all() and (got < n).all(), "row out of range"
    assert m[got].all(), "result row does not pass the filter"
    got_sims_oracle = sim_all[got]
    for i in range(count_got):
        a, b = got_sims_oracle[i], want_sims[i]
        if math.isnan(b):
            assert math.isnan(a), f"pos {i}: oracle NaN, got row {got[i]} sim {a}"
            assert math.isnan(scores_got[i]), f"pos {i}: score should be NaN"
            continue
        assert not math.isnan(a), f"pos {i}: got NaN row before NaN section"
        assert abs(a - b) <= tie_tol, f"pos {i}: row {got[i]} oracle sim {a} vs oracle rank-{i} sim {b}"
        lo_sep = i == 0 or abs(want_sims[i - 1] - b) > 2 * tie_tol or math.isnan(want_sims[i - 1])
        hi_sep = i == len(ext_sims) - 1 or math.isnan(ext_sims[i + 1]) or abs(ext_sims[i + 1] - b) > 2 * tie_tol
        if lo_sep and hi_sep:
            assert got[i] == want_rows[i], f"pos {i}: id {got[i]} != oracle {want_rows[i]} (no tie)"
        tol = rtol * max(abs(a), atol_floor)
        assert abs(float(scores_got[i]) - a) <= tol, f"pos {i}: score {scores_got[i]} vs oracle {a} (tol {tol})"
    # beyond count: padded
    for i in range(count_got, len(rows_got)):
        assert rows_got[i] == -1, f"pos {i}: padding row should be -1"
