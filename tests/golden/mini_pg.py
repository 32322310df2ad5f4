"""GENERATION-TIME TEST INFRASTRUCTURE (used only by tests/golden/make_golden.py).

A tiny interpreter for the ONE statement shape the reference emits on this path
(vector_store.py:274-287, corpus_search.py:1525-1536):

    SELECT <cols>, 1 - (embedding_vec <=> CAST(:query_vec AS vector)) AS similarity
    FROM <table> [LEFT JOIN document_tags dt ON ...]
    WHERE <boolean formula over =, = ANY(..), ILIKE, jsonb_exists, IS NOT NULL, AND, OR, (..)>
    ORDER BY embedding_vec <=> CAST(:query_vec AS vector) LIMIT :k

It reads the SQL TEXT the reference's own Python produced (so the clause assembly, parameter
naming and tag-mode logic under test are the reference's, not a restatement), evaluates the WHERE
per row on string columns with Postgres semantics (NULL comparisons are not true; the formula has
no NOT, so "NULL = false" is exact), and orders by the pgvector distance restatement
(oracle/pgv_oracle.c; NaN last, ties in heap order = row order).
"""
from __future__ import annotations

import re

import numpy as np


def _ilike(value, pattern) -> bool:
    if value is None or pattern is None:
        return False
    rx = "".join(".*" if c == "%" else "." if c == "_" else re.escape(c) for c in pattern)
    return re.fullmatch(rx, value, flags=re.IGNORECASE | re.DOTALL) is not None


_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lpar>\() | (?P<rpar>\)) | (?P<comma>,)
  | (?P<str>'(?:[^']|'')*')
  | (?P<param>:[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op><=>|=|\[\]|\?|\*)
  | (?P<num>\d+)
  | (?P<word>[A-Za-z_][A-Za-z_0-9.]*)
""", re.X)


def _tokens(s: str):
    pos, out = 0, []
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m:
            raise ValueError(f"mini_pg: cannot tokenise at {s[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group()))
    return out


class _Where:
    """Recursive-descent parser -> closure(row_dict) -> bool."""

    def __init__(self, sql: str, params: dict):
        self.t, self.i, self.params = _tokens(sql), 0, params

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def take(self, value=None):
        kind, v = self.peek()
        if value is not None and v.upper() != value:
            raise ValueError(f"mini_pg: expected {value}, got {v!r}")
        self.i += 1
        return kind, v

    def parse(self):
        f = self.expr_or()
        if self.peek()[0] != "eof":
            raise ValueError(f"mini_pg: trailing tokens {self.t[self.i:self.i + 5]}")
        return f

    def expr_or(self):
        fs = [self.expr_and()]
        while self.peek()[1].upper() == "OR":
            self.take()
            fs.append(self.expr_and())
        return fs[0] if len(fs) == 1 else (lambda r, fs=fs: any(f(r) for f in fs))

    def expr_and(self):
        fs = [self.atom()]
        while self.peek()[1].upper() == "AND":
            self.take()
            fs.append(self.atom())
        return fs[0] if len(fs) == 1 else (lambda r, fs=fs: all(f(r) for f in fs))

    def value(self):
        kind, v = self.take()
        if kind == "param":
            return self.params[v[1:]]
        if kind == "str":
            return v[1:-1].replace("''", "'")
        raise ValueError(f"mini_pg: expected a value, got {v!r}")

    def atom(self):
        kind, v = self.peek()
        if kind == "lpar":
            self.take()
            f = self.expr_or()
            self.take(")")
            return f
        if kind == "word" and v.lower() == "jsonb_exists":
            self.take(); self.take("(")
            col = self.take()[1]
            self.take(",")
            key = self.value()
            self.take(")")
            name = {"dt.d_tags": "_doc_d_tags", "dt.p_tags": "_doc_p_tags"}[col.lower()]
            # LEFT JOIN: no document_tags row -> dt.* is NULL -> jsonb_exists(NULL, k) is NULL
            return lambda r, name=name, key=key: r[name] is not None and key in r[name]
        if kind == "num":                                   # WHERE 1=1
            a = self.take()[1]; self.take("="); b = self.take()[1]
            return lambda r, a=a, b=b: a == b
        if kind != "word":
            raise ValueError(f"mini_pg: unexpected token {v!r}")
        col = self.take()[1].split(".")[-1]
        nk, nv = self.peek()
        if nv == "?":                                       # jsonb ? key  (NULL ? key is NULL)
            self.take()
            key = self.value()
            return lambda r, col=col, key=key: r[col] is not None and key in r[col]
        if nv.upper() == "IS":
            self.take(); self.take("NOT"); self.take("NULL")
            return lambda r, col=col: r[col] is not None
        if nv.upper() == "ILIKE":
            self.take()
            pat = self.value()
            return lambda r, col=col, pat=pat: _ilike(r[col], pat)
        if nv == "=":
            self.take()
            if self.peek()[1].upper() == "ANY":
                self.take(); self.take("(")
                if self.peek()[1].upper() == "CAST":
                    self.take(); self.take("(")
                    arr = self.value()
                    self.take("AS"); self.take()            # uuid
                    if self.peek()[1] == "[]":
                        self.take()
                    self.take(")")
                else:
                    arr = self.value()
                self.take(")")
                s = set(arr)
                return lambda r, col=col, s=s: r[col] is not None and r[col] in s
            val = self.value()
            return lambda r, col=col, val=val: r[col] is not None and val is not None and r[col] == val
        raise ValueError(f"mini_pg: unsupported predicate after {col!r}: {nv!r}")


_STMT = re.compile(
    r"SELECT(?P<cols>.*?)FROM\s+(?P<table>[A-Za-z_]+)\s*(?P<join>LEFT\s+JOIN\s+document_tags\s+dt\s+ON\s+[^\n]*?)?\s*"
    r"WHERE(?P<where>.*?)ORDER\s+BY\s+embedding_vec\s*<=>\s*CAST\(:query_vec\s+AS\s+vector\)\s*LIMIT\s+:k\s*$",
    re.S | re.I)


def parse_vector_text(s: str) -> np.ndarray:
    """'[f1,f2,...]' -> float4 array, as pgvector's vector_in (strtof per element)."""
    body = s.strip()
    assert body[0] == "[" and body[-1] == "]"
    return np.asarray([np.float32(float(x)) for x in body[1:-1].split(",")], dtype=np.float32)


def execute(table_rows: list[dict], X: np.ndarray, sql: str, params: dict, cosine_distance) -> list[dict]:
    """Run the statement.  table_rows[i] holds the string columns of row i plus '_doc_d_tags' /
    '_doc_p_tags' (set or None = no document_tags row) and 'embedding_vec' (None = SQL NULL);
    X[i] is its float4 vector; cosine_distance(X, q) -> float8 distances (NaN allowed)."""
    m = _STMT.search(sql.strip())
    if not m:
        raise ValueError("mini_pg: statement shape not recognised:\n" + sql)
    where = _Where(m.group("where"), params).parse()
    q = parse_vector_text(params["query_vec"])
    passing = [i for i, r in enumerate(table_rows) if where(r)]
    if not passing:
        return []
    d = np.asarray(cosine_distance(X[passing], q), dtype=np.float64)
    nan = np.isnan(d)
    order = np.lexsort((np.arange(len(passing)), np.where(nan, np.inf, d), nan))   # NaN last, ties by row
    out = []
    for j in order[: int(params["k"])]:
        r = dict(table_rows[passing[j]])
        sim = 1.0 - d[j]
        r["similarity"] = float(sim)
        out.append(r)
    return out


_DTAG_STMT = re.compile(
    r"SELECT(?P<cols>.*?)0\.5\s+AS\s+similarity\s+FROM\s+rag_published_embeddings\s+WHERE(?P<where>.*?)"
    r"ORDER\s+BY\s+CASE\s+document_authority_level\s+WHEN\s+'contract_source_of_truth'\s+THEN\s+0\s+"
    r"WHEN\s+'operational'\s+THEN\s+1\s+ELSE\s+2\s+END\s*,\s*rag_published_embeddings\.id\s+LIMIT\s+:k\s*$", re.S | re.I)
_COUNT_STMT = re.compile(r"SELECT\s+COUNT\(\*\)\s+AS\s+n_total\s*,(?P<filters>.*?)FROM\s+rag_published_embeddings\s+WHERE(?P<where>.*)$",
                         re.S | re.I)
_COUNT_ITEM = re.compile(r"COUNT\(\*\)\s+FILTER\s+\(WHERE\s+(?P<cond>.*?)\)\s+AS\s+(?P<name>cnt_\d+)", re.S | re.I)


def execute_any(table_rows: list[dict], X, sql: str, params: dict, cosine_distance) -> list[dict]:
    """Dispatch on the three statement shapes of the path: the vector statement, the d-tag arm's SELECT
    (corpus_search.py:1664-1681) and its IDF COUNT (corpus_search.py:1644-1648)."""
    text = sql.strip()
    if "<=>" in text:
        return execute(table_rows, X, sql, params, cosine_distance)
    m = _COUNT_STMT.search(text)
    if m and "COUNT(*)" in text.upper():
        where = _Where(m.group("where"), params).parse()
        passing = [r for r in table_rows if where(r)]
        out = {"n_total": len(passing)}
        for it in _COUNT_ITEM.finditer(m.group("filters")):
            cond = _Where(it.group("cond"), params).parse()
            out[it.group("name")] = sum(1 for r in passing if cond(r))
        return [out]
    m = _DTAG_STMT.search(text)
    if not m:
        raise ValueError("mini_pg: statement shape not recognised:\n" + sql)
    where = _Where(m.group("where"), params).parse()
    passing = [r for r in table_rows if where(r)]

    def tier(r):
        lvl = r["document_authority_level"]
        return 0 if lvl == "contract_source_of_truth" else 1 if lvl == "operational" else 2
    passing.sort(key=lambda r: (tier(r), r["id"]))          # uuid order == order of the canonical lowercase text
    return [dict(r, similarity=0.5) for r in passing[: int(params["k"])]]
