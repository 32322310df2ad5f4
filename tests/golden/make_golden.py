#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by running the REFERENCE'S OWN PYTHON for the
hot path, imported from /root/reference (build container only; the GPU box never needs it).

    python tests/golden/make_golden.py

What runs here is the reference's code, unmodified:
  * app/services/corpus_search.py  `_vector_arm` (1427-1602) incl. `_build_filter_clauses`
    (516-560), tag-mode logic (1464-1523), retry (1538-1551), post-processing (1565-1602),
    `_row_to_base_dict` (563-587);
  * app/services/vector_store.py   `PgVectorStore._search_async` (228-303).
Only two things are substituted, because they do not exist in this container:
  * `sqlalchemy.text` / `AsyncSession` -> 10-line stand-ins (the reference only wraps a string);
  * the Postgres server -> tests/golden/mini_pg.py, an interpreter for the one statement shape the
    reference emits, reading the SQL text the reference produced; `<=>` is the pgvector
    restatement of oracle/pgv_oracle.c (that arithmetic stays "parity unpinned").
The fixtures pin everything the reference's Python decides: which clauses are emitted for which
inputs, parameter values, strict/relaxed/auto behaviour, LIMIT, clamp, min_similarity, stop-at-k,
dict shape.  tests/test_golden.py replays them against the oracle (CPU) and the product (GPU).
"""
from __future__ import annotations

import asyncio
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)


def install_stubs():
    sa = types.ModuleType("sqlalchemy")

    class _Text:
        def __init__(self, s):
            self.text = s

        def __str__(self):
            return self.text

    sa.text = lambda s: _Text(s)
    ext = types.ModuleType("sqlalchemy.ext")
    aio = types.ModuleType("sqlalchemy.ext.asyncio")

    class AsyncSession:  # annotation only
        pass

    aio.AsyncSession = AsyncSession
    sys.modules.update({"sqlalchemy": sa, "sqlalchemy.ext": ext, "sqlalchemy.ext.asyncio": aio})


class FakeResult:
    def __init__(self, rows):
        self._rows = rows

    def mappings(self):
        return self

    def all(self):
        return self._rows


class FakeSession:
    """Stands in for AsyncSession / AsyncSessionLocal(): records every statement, answers through mini_pg."""

    def __init__(self, table_rows, X, cosine_distance, log):
        self.table_rows, self.X, self.cd, self.log = table_rows, X, cosine_distance, log

    async def execute(self, stmt, params):
        import mini_pg
        sql = str(stmt)
        self.log.append({"sql": " ".join(sql.split()),
                         "params": {k: v for k, v in params.items() if k != "query_vec"}})
        return FakeResult(mini_pg.execute(self.table_rows, self.X, sql, dict(params), self.cd))

    async def __aenter__(self):
        return self

    async def __aexit__(self, *a):
        return False


def main():
    if not os.path.isdir(REF):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    install_stubs()
    sys.path.insert(0, REF)
    from oracle import oracle
    from helpers import build_tables
    import importlib
    ref_cs = importlib.import_module("app.services.corpus_search")
    ref_lex = importlib.import_module("app.services.corpus_search_lexicon")
    ref_vs = importlib.import_module("app.services.vector_store")

    n, dim = 1500, 32
    ot, _, X, valid, meta, info = build_tables(oracle, n, dim, seed=11, dtype="f32", null_frac=4e-3,
                                               rows_per_doc=12, with_product=False)
    # ---- the table as Postgres would hold it
    table_rows = []
    for i in range(n):
        did = ot.document_id[i]
        r = {"id": ot.id[i], "document_id": did, "source_type": ot.source_type[i], "source_id": ot.source_id[i],
             "document_payer": ot.document_payer[i], "document_state": ot.document_state[i],
             "document_program": ot.document_program[i], "document_authority_level": ot.document_authority_level[i],
             "embedding_vec": (True if ot.has_vec[i] else None),
             "_doc_d_tags": ot.doc_d_tags.get(did), "_doc_p_tags": ot.doc_p_tags.get(did)}
        for c, col in ot.extra.items():
            r[c] = col[i]
        table_rows.append(r)
    Xf = np.ascontiguousarray(ot.X, dtype=np.float32)

    def cd(Xs, q):
        with np.errstate(all="ignore"):
            return oracle.cosine_distance_c(np.ascontiguousarray(Xs), q)

    rng = np.random.default_rng(2024)

    def emb(i):
        if i % 2 == 0:
            return (X[int(rng.integers(0, n))].astype(np.float64) + 0.05 * rng.standard_normal(dim)).tolist()
        return rng.standard_normal(dim).tolist()

    CF = ref_cs.CorpusFilters
    E = ref_lex.LexiconExpansion
    docs = sorted(set(ot.document_id))
    pool19 = [docs[int(j)] for j in rng.choice(len(docs), 19, replace=False)]
    pool_big = [docs[int(j)] for j in rng.choice(len(docs), min(100, len(docs)), replace=False)]
    zero_rows = np.nonzero((np.abs(X).sum(axis=1) == 0) & valid.astype(bool))[0]
    cases = [
        dict(k=10),
        dict(k=20, over_fetch_factor=8, min_similarity=0.3),
        dict(k=200, over_fetch_factor=8),
        dict(k=10, filters=dict(payer="Sunshine Health")),
        dict(k=10, filters=dict(payer="Centene", state="FL")),
        dict(k=10, filters=dict(program="Medicaid", authority_level="payer_policy")),
        dict(k=10, filters=dict(payer="Molina Healthcare", state="FL", program="Medicaid")),
        dict(k=10, include_document_ids=pool19),
        dict(k=40, include_document_ids=pool_big, filters=dict(state="FL")),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:state.fl", "j:bad"]), tag_mode="auto"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:payor.molina_healthcare", "j:program.medic"]), tag_mode="strict"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:regulatory_authority.ahca"], domain_tags=["d:topic_000.leaf"]), tag_mode="auto"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"], process_tags=["p:topic_003.leaf"]), tag_mode="auto"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"]), tag_mode="strict"),
        dict(k=10, expansion=dict(domain_tags=["d:topic_006.leaf", "d:unknown.key"], process_tags=["p:topic_001.leaf"]), tag_mode="relaxed"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:state.fl"], domain_tags=["d:topic_000.leaf"]), tag_mode="none"),
        dict(k=10, expansion=dict(domain_tags=["d:topic_000.leaf"]), tag_mode="auto"),
        dict(k=10, expansion=dict(jurisdiction_tags=["j:program.medicare_advantage", "j:payor.sunshine"],
                                  domain_tags=["d:topic_002.leaf"]), tag_mode="auto",
             filters=dict(state="FL"), include_document_ids=pool_big),
        dict(k=5, min_similarity=0.999),
        dict(k=3, min_similarity=0.2, over_fetch_factor=8),
    ]
    if len(zero_rows):
        z = int(zero_rows[0])
        cases.append(dict(k=100, include_document_ids=[ot.document_id[z]], _note="NaN similarity row reports 1.0 (corpus_search.py:1569)"))

    arm_out = []
    for i, kw in enumerate(cases):
        q = emb(i)
        log = []
        db = FakeSession(table_rows, Xf, cd, log)
        filters = CF(**kw["filters"]) if kw.get("filters") else None
        expansion = E(**kw["expansion"]) if kw.get("expansion") else None
        got = asyncio.run(ref_cs._vector_arm(
            db, q, kw["k"], filters, kw.get("include_document_ids"), search_id="",
            expansion=expansion, tag_mode=kw.get("tag_mode", "auto"),
            min_similarity=kw.get("min_similarity"), over_fetch_factor=kw.get("over_fetch_factor", 1)))
        arm_out.append({"case": {k: v for k, v in kw.items()}, "query": q, "statements": log, "result": got})

    # ---- PgVectorStore._search_async
    dbmod = types.ModuleType("app.database")
    store_log = []
    dbmod.AsyncSessionLocal = lambda: FakeSession(table_rows, Xf, cd, store_log)
    sys.modules["app.database"] = dbmod
    store = ref_vs.PgVectorStore()
    doc = ot.document_id[777]
    store_cases = [
        dict(k=10), dict(k=1), dict(k=100),
        dict(k=10, document_id=doc),
        dict(k=10, filters={"payer": "Sunshine Health"}),
        dict(k=10, filters={"payer": "Sunshine Health", "state": "FL", "authority_level": "payer_policy"}),
        dict(k=10, filters={"state": "", "payer": None, "bogus": "x", "source_type": "fact"}),
        dict(k=10, filters={"document_id": doc, "source_type": "hierarchical"}),
        dict(k=10, filters={"payer": "No Such Payer"}),
    ]
    store_out = []
    for i, kw in enumerate(store_cases):
        q = emb(i + 100)
        del store_log[:]
        got = asyncio.run(store.asearch(q, kw["k"], kw.get("document_id"), kw.get("filters")))
        store_out.append({"case": kw, "query": q, "statements": list(store_log), "result": got})

    # ---- write fixtures
    np.savez_compressed(os.path.join(HERE, "table_vectors.npz"), X=Xf, has_vec=np.asarray(ot.has_vec, dtype=np.uint8))
    table_json = {
        "n": n, "dim": dim,
        "columns": {c: getattr(ot, c) for c in ("id", "document_id", "source_type", "source_id", "document_payer",
                                                "document_state", "document_program", "document_authority_level")},
        "extra": ot.extra,
        "doc_d_tags": {k: sorted(v) for k, v in ot.doc_d_tags.items()},
        "doc_p_tags": {k: sorted(v) for k, v in ot.doc_p_tags.items()},
    }
    json.dump(table_json, open(os.path.join(HERE, "table.json"), "w"), separators=(",", ":"))
    json.dump(arm_out, open(os.path.join(HERE, "vector_arm.json"), "w"), separators=(",", ":"))
    json.dump(store_out, open(os.path.join(HERE, "store_search.json"), "w"), separators=(",", ":"))
    print(f"wrote {len(arm_out)} _vector_arm cases, {len(store_out)} PgVectorStore cases; "
          f"result sizes {[len(c['result']) for c in arm_out]}")
    print("statements per case:", [len(c["statements"]) for c in arm_out])


if __name__ == "__main__":
    main()
