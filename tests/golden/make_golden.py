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
import gzip
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


def dump_json(obj, name):
    """Fixtures are gzip-compressed JSON (mtime 0 so regeneration is byte-stable)."""
    with gzip.GzipFile(os.path.join(HERE, name + ".gz"), "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode())


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

    def one_or_none(self):
        return self._rows[0] if self._rows else None


class FakeSession:
    """Stands in for AsyncSession / AsyncSessionLocal(): records every statement, answers through mini_pg."""

    def __init__(self, table_rows, X, cosine_distance, log):
        self.table_rows, self.X, self.cd, self.log = table_rows, X, cosine_distance, log

    async def execute(self, stmt, params):
        import mini_pg
        sql = str(stmt)
        self.log.append({"sql": " ".join(sql.split()),
                         "params": {k: v for k, v in params.items() if k != "query_vec"}})
        return FakeResult(mini_pg.execute_any(self.table_rows, self.X, sql, dict(params), self.cd))

    async def __aenter__(self):
        return self

    async def __aexit__(self, *a):
        return False


PHRASE_POOL = [
    "prior authorization", "sunshine health", "timely filing", "appeal", "medical records", "molina healthcare",
    "behavioral health", "provider services", "claims submission", "credentialing", "telehealth", "phone",
    "peer support", "targeted case management", "florida medicaid", "denial",
]
FILLER = ("the plan requires that providers follow the documented process for each covered service and retain "
          "supporting notes for review by the health plan within the stated period of time").split()


def hybrid_table(n=900, dim=32, seed=5):
    """A small text-rich table for the `_rerank` fixtures: bodies assembled from filler words and pool
    phrases, varied lengths, some phone numbers, chunk d-tags, document d/j/p tags."""
    rng = np.random.default_rng(seed)
    n_docs = 60
    doc_of = np.sort(rng.integers(0, n_docs, size=n))
    payers = ["Sunshine Health", "Molina Healthcare", "AHCA", "Aetna", None]
    auths = ["contract_source_of_truth", "payer_website", "operational_suggested", "payer_policy", "fyi_not_citable", None, "Mystery"]
    dkeys = ["utilization_management.prior_authorization", "claims.timely_filing", "claims.appeals", "benefits.behavioral_health"]
    jkeys = ["payor.sunshine_health", "payor.molina_healthcare", "state.fl", "program.medicaid"]
    docs = []
    for d in range(n_docs):
        payer = payers[int(rng.integers(0, len(payers)))]
        docs.append({
            "document_id": f"00000000-0000-4000-8000-{d:012d}",
            "document_payer": payer, "document_state": ["FL", "TX", None][int(rng.integers(0, 3))],
            "document_program": ["Medicaid", "Medicare Advantage", None][int(rng.integers(0, 3))],
            "document_authority_level": auths[int(rng.integers(0, len(auths)))],
            "document_display_name": ["", f"{payer or 'General'} Provider Manual {d}", "Timely Filing and Appeals Guide"][int(rng.integers(0, 3))],
            "document_filename": [f"FL_SunshineHealth_Caid_PAReq_{d}.pdf", f"molina-healthcare_behavioral-health_{d}.pdf", f"doc_{d}.pdf"][int(rng.integers(0, 3))],
            "d_tags": [k for k in dkeys if rng.random() < 0.25], "j_tags": [k for k in jkeys if rng.random() < 0.3],
            "p_tags": ["process.claims_submission"] if rng.random() < 0.2 else [],
            "has_tags_row": bool(rng.random() < 0.8),
        })
    rows = []
    X = rng.standard_normal((n, dim)).astype(np.float32) * np.exp(0.25 * rng.standard_normal((n, 1))).astype(np.float32)
    X[7] = 0.0                                            # zero-norm row: similarity NaN -> reports 1.0
    X[40:46] = X[39]                                      # duplicate cluster
    for i in range(n):
        d = docs[int(doc_of[i])]
        words = []
        for _ in range(int(rng.integers(1, 90))):
            if rng.random() < 0.12:
                words.append(PHRASE_POOL[int(rng.integers(0, len(PHRASE_POOL)))])
            else:
                words.append(FILLER[int(rng.integers(0, len(FILLER)))])
        if rng.random() < 0.05:
            words.append(["call 1-800-555-1234", "fax (305) 555-0100", "dial 813.555.0199"][int(rng.integers(0, 3))])
        text = " ".join(words)
        if rng.random() < 0.1:
            text = text.upper()
        cd = {}
        if rng.random() < 0.08:
            cd[dkeys[int(rng.integers(0, len(dkeys)))]] = 1
        rows.append({
            "id": f"11111111-0000-4000-8000-{i:012d}", "document_id": d["document_id"],
            "source_type": ["hierarchical", "fact", None][int(rng.integers(0, 3))], "source_id": f"src-{i}",
            "text": text if rng.random() > 0.01 else None,
            "page_number": int(i % 30) + 1, "paragraph_index": int(i % 9),
            "section_path": ["", "Claims/Timely_Filing", "Utilization-Management/Prior_Authorization", None][int(rng.integers(0, 4))],
            "chapter_path": [None, "Provider.Services/Contact"][int(rng.integers(0, 2))],
            "summary": [None, "how to appeal a denial"][int(rng.random() < 0.1)],
            "content_sha": f"{i:040x}",
            "document_display_name": d["document_display_name"], "document_filename": d["document_filename"],
            "document_authority_level": d["document_authority_level"], "document_payer": d["document_payer"],
            "document_state": d["document_state"], "document_program": d["document_program"],
            "chunk_d_tags": cd or None, "chunk_p_tags": None, "chunk_j_tags": None,
            "has_vec": bool(rng.random() > 0.01),
        })
    promoted = sorted(int(x) for x in rng.choice(n, 6, replace=False))
    return docs, rows, X, promoted


def make_rerank_golden(oracle, ref_cs):
    """Runs the reference's own `_rerank` over ALL rows of a small table (vector arm only)."""
    cfg = types.ModuleType("app.config")
    cfg.CHUNK_TAG_BOOST = 1.5                         # app/config.py:129 (the module itself needs DATABASE_URL + dotenv)
    sys.modules["app.config"] = cfg
    docs, rows, X, promoted = hybrid_table()
    doc_by_id = {d["document_id"]: d for d in docs}
    rng = np.random.default_rng(77)
    dim = X.shape[1]
    cases = [
        dict(query="sunshine health prior authorization requirements", phrases=["Sunshine Health", "prior authorization"],
             weights=[0.93, 0.79], codes=["j:payor.sunshine_health", "d:utilization_management.prior_authorization"]),
        dict(query="what is the appeal process for a denial", phrases=None, weights=None, codes=None),
        dict(query="sunshine health provider services phone number", phrases=["sunshine health", "phone"],
             weights=[0.9, 0.7], codes=["j:payor.sunshine_health", None]),
        dict(query="timely filing deadline", phrases=["timely filing"], weights=[0.8], codes=["d:claims.timely_filing"]),
        dict(query="timely filing deadline for corrected claim", phrases=["timely filing", "appeal"], weights=None, codes=None),
        dict(query="unicorn coverage", phrases=["unicorn rides"], weights=[1.0], codes=[None]),
        dict(query="", phrases=["molina healthcare", "behavioral health", "credentialing"], weights=[0.9, 0.8, 0.0],
             codes=["j:payor.molina_healthcare", "d:benefits.behavioral_health", None]),
        dict(query="telehealth medical records documentation required", phrases=["medical records"], weights=[0.0], codes=[None]),
        dict(query="prior authorization", phrases=["prior authorization", "florida medicaid"], weights=[0.7, 0.95],
             codes=["d:utilization_management.prior_authorization", "j:state.fl"], payer="AHCA"),   # not an FL-MCO payer: plain equality (:524-535)
        dict(query="peer support claims submission", phrases=["peer support", "claims submission", "denial", "appeal"],
             weights=[0.9, 0.8, 0.7, 0.6], codes=[None, "p:process.claims_submission", None, "d:claims.appeals"]),
    ]
    out_cases = []
    for ci, kw in enumerate(cases):
        q = (X[int(rng.integers(0, len(rows)))].astype(np.float64) + 0.1 * rng.standard_normal(dim)).tolist() if ci % 2 == 0 \
            else rng.standard_normal(dim).tolist()
        qv = np.asarray([np.float32(x) for x in q], dtype=np.float32)
        with np.errstate(all="ignore"):
            dist = oracle.cosine_distance_c(np.ascontiguousarray(X), qv)
        cands = []
        for i, r in enumerate(rows):
            if not r["has_vec"]:
                continue
            if kw.get("payer") and r["document_payer"] != kw["payer"]:
                continue
            sim = 1.0 - float(dist[i])
            cosine_sim = max(0.0, min(1.0, float(sim or 0.0)))           # _vector_arm, corpus_search.py:1569
            c = ref_cs._row_to_base_dict(r)
            c["similarity"] = cosine_sim
            c["match_score"] = cosine_sim
            c["_arm"] = "vector"
            c["arm_scores"] = {"vector": cosine_sim}                     # _rrf_merge keeps the raw per-arm score
            c["retrieval_arms"] = ["vector"]
            d = doc_by_id[r["document_id"]]
            if d["has_tags_row"]:                                        # _attach_inherited_doc_tags, :2802-2807
                c["_doc_d_tags"] = list(d["d_tags"]); c["_doc_j_tags"] = list(d["j_tags"]); c["_doc_p_tags"] = list(d["p_tags"])
            if i in promoted:
                c["_promoted_from_seed"] = "seed"
            cands.append(c)
        ranked = ref_cs._rerank(cands, "", kw["query"], kw["phrases"], kw["weights"], kw["codes"])
        out_cases.append({
            "case": kw, "query_embedding": q, "n_candidates": len(cands), "n_ranked": len(ranked),
            "top": [{"id": c["id"], "rerank_score": c["rerank_score"], "similarity": c["similarity"],
                     "source_type": c["source_type"]} for c in ranked[:50]],
        })
    # ---- `_dtag_arm` (corpus_search.py:1605-1701) and `_rrf_merge` (:1708-1766) on the same table
    table_rows = []
    for r in rows:
        tr = dict(r)
        tr["embedding_vec"] = True if r["has_vec"] else None
        table_rows.append(tr)
    CF = ref_cs.CorpusFilters
    dtag_cases = [
        dict(keys=["claims.timely_filing"], k=10),
        dict(keys=["claims.timely_filing", "claims.appeals", "no.such_key"], k=25, idf_mode=True),
        dict(keys=["utilization_management.prior_authorization"], k=100, filters=dict(payer="AHCA"), idf_mode=True),
        dict(keys=["benefits.behavioral_health"], k=5, include_document_ids=[docs[i]["document_id"] for i in range(0, 60, 3)], idf_mode=True),
        dict(keys=["no.such_key"], k=10, idf_mode=True),
        dict(keys=[], k=10),
    ]
    dtag_out = []
    for kw in dtag_cases:
        log = []
        db = FakeSession(table_rows, X, None, log)
        got = asyncio.run(ref_cs._dtag_arm(db, kw["keys"], kw["k"], CF(**kw["filters"]) if kw.get("filters") else None,
                                           kw.get("include_document_ids"), "", kw.get("idf_mode", False)))
        dtag_out.append({"case": kw, "statements": log, "result": got})
    # RRF over three arms built from real outputs: vector order, a shuffled "bm25" list, the dtag arm with IDF
    vec = [dict(c) for c in out_cases and []]
    qv = np.asarray([np.float32(x) for x in out_cases[0]["query_embedding"]], dtype=np.float32)
    with np.errstate(all="ignore"):
        dist = oracle.cosine_distance_c(np.ascontiguousarray(X), qv)
    order = [i for i in np.argsort(dist, kind="stable") if rows[i]["has_vec"] and not np.isnan(dist[i])][:40]
    vec_arm = []
    for i in order:
        c = ref_cs._row_to_base_dict(rows[i]); c["similarity"] = 1.0 - float(dist[i]); c["_arm"] = "vector"; vec_arm.append(c)
    bm_idx = [int(x) for x in rng.permutation(len(rows))[:30]] + order[5:12]
    bm_arm = []
    for j, i in enumerate(bm_idx):
        c = ref_cs._row_to_base_dict(rows[i]); c["similarity"] = 0.9 - 0.01 * j; c["_arm"] = "bm25"
        if j % 4 == 0:
            c["summary"] = "bm25 summary"             # fills a blank of an earlier arm's dict
        bm_arm.append(c)
    arms = {"bm25": bm_arm, "vector": vec_arm, "dtag": dtag_out[1]["result"]}
    fused = ref_cs._rrf_merge({a: [dict(c) for c in lst] for a, lst in arms.items()})
    dump_json({"dtag": dtag_out, "rrf": {"arms": arms, "fused": fused}}, "dtag_rrf.json")
    # ---- the text rules the product's shim restates (hybrid.py), evaluated by the reference's own helpers
    helper_rows = []
    for i in list(range(0, len(rows), 9))[:100]:
        c = ref_cs._row_to_base_dict(rows[i])
        d = doc_by_id[rows[i]["document_id"]]
        if d["has_tags_row"]:
            c["_doc_d_tags"] = list(d["d_tags"]); c["_doc_j_tags"] = list(d["j_tags"]); c["_doc_p_tags"] = list(d["p_tags"])
        body = ref_cs._body_haystack(c)
        helper_rows.append({
            "row": i, "body_haystack": body, "meta_haystack": ref_cs._meta_haystack(c),
            "classify_jpd": ref_cs._classify_jpd(body), "length_score": ref_cs._length_score(c.get("text") or ""),
            "authority_score": ref_cs._authority_score(c.get("authority_level")),
            "contact_value": bool(ref_cs._CONTACT_VALUE_RE.search(c.get("text") or "")),
        })
    queries = ["sunshine health provider services phone number", "what is the appeal process", "EDI payer-id for claims",
               "toll free hotline", "prior authorization criteria for inpatient level of care", ""]
    helper_q = [{"query": q, "classify_jpd": ref_cs._classify_jpd(q) if q else {}, "contact_query": bool(ref_cs._CONTACT_QUERY_RE.search(q)) if q else False}
                for q in queries]
    dump_json({"rows": helper_rows, "queries": helper_q, "confidence": {str(x): ref_cs._confidence_label(x) for x in (0.0, 0.17, 0.18, 0.34, 0.35, 0.54, 0.55, 0.9)}},
              "text_rules.json")
    print("dtag results:", [len(c["result"]) for c in dtag_out], "rrf fused:", len(fused))
    np.savez_compressed(os.path.join(HERE, "hybrid_vectors.npz"), X=X)
    dump_json({"docs": docs, "rows": rows, "promoted": promoted, "phrase_pool": PHRASE_POOL}, "hybrid_table.json")
    dump_json(out_cases, "rerank.json")
    print("rerank cases:", [(c["n_candidates"], c["n_ranked"]) for c in out_cases])


def make_pool_golden():
    """Runs the reference's own `build_candidate_pool` + `_augment_pool_with_inheritance`
    (app/services/corpus_search_agent.py:1762-1888, 1966-2002) over a synthetic document_tags table; the only stand-in is
    the session that answers `SELECT document_id FROM document_tags WHERE {column} ? :code` (:1461-1482)."""
    import importlib
    import re
    os.environ.setdefault("DATABASE_URL", "postgresql+asyncpg://u:p@localhost/db")     # app/config.py refuses to import without one
    agent = importlib.import_module("app.services.corpus_search_agent")
    rng = np.random.default_rng(909)
    jkeys = ["payor.sunshine_health", "payor.aetna", "payor.molina_healthcare", "state.fl", "program.medicaid", "regulatory_authority.ahca"]
    dkeys = ["utilization_management.prior_authorization", "claims.timely_filing", "claims.appeals", "benefits.behavioral_health", "claims.general"]
    pkeys = ["process.claims_submission", "process.appeal_filing"]
    docs = []
    for d in range(400):
        docs.append({
            "document_id": f"00000000-0000-4000-8000-{d:012d}",
            "j_tags": [k for k, p in zip(jkeys, (0.2, 0.15, 0.1, 0.5, 0.4, 0.12)) if rng.random() < p],
            "d_tags": [k for k, p in zip(dkeys, (0.15, 0.1, 0.1, 0.05, 0.4)) if rng.random() < p],
            "p_tags": [k for k, p in zip(pkeys, (0.15, 0.05)) if rng.random() < p],
            "has_tags_row": bool(rng.random() < 0.9),
        })
    stmt = re.compile(r"SELECT document_id FROM document_tags WHERE (j_tags|d_tags|p_tags) \? :code")

    class TagSession:
        def __init__(self):
            self.log = []

        async def execute(self, sql, params):
            m = stmt.fullmatch(" ".join(str(sql).split()))
            assert m, str(sql)
            self.log.append([m.group(1), params["code"]])
            rows = [(d["document_id"],) for d in docs if d["has_tags_row"] and params["code"] in d[m.group(1)]]

            class R:
                def all(self_inner):
                    return rows
            return R()

    TA, TP = agent.TermAssignment, agent.TermPartition
    import dataclasses
    ta_fields = {f.name for f in dataclasses.fields(TA)}

    def term(code):
        kw = {"kind": "tag", "full_code": code}
        for name in ta_fields - set(kw):
            kw[name] = {"term": code, "selectivity": 0.9}.get(name, None)
        return TA(**kw)

    cases = [
        dict(required=["j:payor.sunshine_health", "d:claims.timely_filing", "p:process.claims_submission"]),
        dict(required=["j:payor.sunshine_health", "d:claims.timely_filing"], boosted=["p:process.appeal_filing"]),
        dict(required=["j:payor.sunshine_health", "j:state.fl", "d:utilization_management.prior_authorization"]),
        dict(required=["j:payor.molina_healthcare", "d:benefits.behavioral_health", "d:claims.appeals"]),     # J&D empty -> AHCA&D or AHCA
        dict(required=["d:claims.general"]),
        dict(required=["j:payor.aetna"]),
        dict(required=["j:payor.nobody", "d:claims.general"]),                                                # unknown j code
        dict(required=["d:no.such_domain"]),
        dict(required=[]),
        dict(required=["p:process.claims_submission"], boosted=["d:claims.appeals", "j:program.medicaid"]),
        dict(required=["j:payor.aetna", "d:claims.general"], inherited=[docs[i]["document_id"] for i in (3, 5, 8, 399)] + ["ffffffff-0000-4000-8000-000000000000"]),
    ]
    out = []
    for kw in cases:
        part = TP(required=[term(c) for c in kw.get("required", [])], boosted=[term(c) for c in kw.get("boosted", [])], dropped=[])
        db = TagSession()
        pool = asyncio.run(agent.build_candidate_pool(db, part))
        if kw.get("inherited"):
            pool = agent._augment_pool_with_inheritance(pool, kw["inherited"])
        out.append({"case": kw, "statements": db.log, "cascade_level": pool.cascade_level,
                    "cascade_steps": [list(x) for x in pool.cascade_steps], "intersect_codes": pool.intersect_codes,
                    "document_ids": sorted(pool.document_ids), "inherited_document_ids": list(pool.inherited_document_ids),
                    "relaxed": pool.relaxed})
    dump_json({"docs": docs, "cases": out}, "pool.json")
    print("pool cases:", [(c["cascade_level"], len(c["document_ids"])) for c in out])


def make_corpus_search_golden(oracle):
    """Runs the reference's own `corpus_search` (app/services/corpus_search.py:3280-3826) end to end over the hybrid table.
    Stand-ins: the Postgres session (statement shapes answered below; the vector / d-tag statements go through mini_pg as in
    the other fixtures), `_bm25_arm` (returns a recorded list: Postgres full-text search is outside the path),
    `_embed_with_cache` (returns the case's embedding), `expand_query_via_lexicon`, and the fire-and-forget
    `_persist_search_event`.  Everything between them -- arm orchestration per mode, RRF, content de-duplication,
    neighbour enrichment, inherited tags, topic-block merge, `_rerank`, `_assemble`, neighbour expansion, chunk shaping --
    is the reference's code."""
    import importlib
    import re
    import mini_pg
    cfg = types.ModuleType("app.config")
    cfg.CHUNK_TAG_BOOST = 1.5
    sys.modules["app.config"] = cfg
    ref_cs = importlib.import_module("app.services.corpus_search")
    ref_lex = importlib.import_module("app.services.corpus_search_lexicon")
    docs, rows, X, promoted = hybrid_table()
    # four consecutive chunks share a page here (the shared fixture puts every chunk on its own page, which would leave the
    # same-page topic-block merge without work); tests/test_gpu_corpus_search.py applies the same renumbering
    rows = [dict(r, page_number=(i // 4) % 30 + 1) for i, r in enumerate(rows)]
    doc_by_id = {d["document_id"]: d for d in docs}
    table_rows = []
    for r in rows:
        tr = dict(r)
        tr["embedding_vec"] = True if r["has_vec"] else None
        d = doc_by_id[r["document_id"]]
        tr["_doc_d_tags"] = set(d["d_tags"]) if d["has_tags_row"] else None
        tr["_doc_p_tags"] = set(d["p_tags"]) if d["has_tags_row"] else None
        table_rows.append(tr)

    def cd(Xs, q):
        with np.errstate(all="ignore"):
            return oracle.cosine_distance_c(np.ascontiguousarray(Xs), q)

    class Result:
        def __init__(self, rows, scalar=None):
            self._rows, self._scalar = rows, scalar

        def mappings(self):
            return self

        def all(self):
            return self._rows

        def scalar(self):
            return self._scalar

    class Session:
        def __init__(self):
            self.log = []

        async def execute(self, stmt, params=None):
            sql = " ".join(str(stmt).split())
            params = dict(params or {})
            self.log.append(sql[:60])
            if "<=>" in sql or "0.5 AS similarity" in sql or "COUNT(*) AS n_total" in sql:
                return Result(mini_pg.execute_any(table_rows, X, str(stmt), params, cd))
            if sql.startswith("SELECT count(*) FROM rag_published_embeddings WHERE document_id = ANY"):
                ids = set(params["_vc_ids"])
                return Result([], scalar=sum(1 for r in rows if r["document_id"] in ids))
            if sql.startswith("SELECT DISTINCT ON (m.id)"):
                hit = {}
                for doc, lo, hi, plo, phi, ex in zip(params["doc_ids"], params["para_lo"], params["para_hi"], params["page_lo"],
                                                     params["page_hi"], params["excludes"]):
                    for r in rows:
                        if (r["document_id"] == doc and r["paragraph_index"] is not None and lo <= r["paragraph_index"] <= hi
                                and r["page_number"] is not None and plo <= r["page_number"] <= phi and r["id"] != ex):
                            hit[r["id"]] = r
                return Result([dict(hit[i]) for i in sorted(hit)][:500])
            if "FROM document_tags" in sql:
                ids = set(params["ids"])
                return Result([{"doc_id": d["document_id"], "d_tags": list(d["d_tags"]), "j_tags": list(d["j_tags"]), "p_tags": list(d["p_tags"])}
                               for d in docs if d["has_tags_row"] and d["document_id"] in ids])
            if "FROM payor_inherited_authority" in sql:
                return Result([])
            raise ValueError("corpus_search golden: statement not recognised: " + sql[:200])

    async def no_persist(*a, **k):
        return None
    ref_cs._persist_search_event = no_persist
    rng = np.random.default_rng(4242)
    dim = X.shape[1]

    def bm25_list(seed_rows, n_extra):
        idx = list(seed_rows) + [int(x) for x in rng.permutation(len(rows))[:n_extra]]
        out, seen = [], set()
        for j, i in enumerate(idx):
            if i in seen:
                continue
            seen.add(i)
            c = ref_cs._row_to_base_dict(rows[i])
            c["similarity"] = round(0.95 - 0.02 * j, 4) if j < 40 else 0.1
            c["match_score"] = c["similarity"]
            c["_arm"] = "bm25"
            out.append(c)
        return out

    cases = [
        dict(req=dict(query="what is the appeal process for a denial", k=8, mode="recall")),
        dict(req=dict(query="timely filing deadline", k=5, mode="recall", min_similarity=0.2, neighbor_paragraph_window=0)),
        dict(req=dict(query="prior authorization requirements", k=10, mode="recall", filters=dict(payer="AHCA")),
             lexicon=dict(domain_tags=["d:claims.timely_filing"], jurisdiction_tags=["j:state.fl"])),
        dict(req=dict(query="sunshine health prior authorization requirements", k=10, mode="corpus",
                      required_phrases=["Sunshine Health", "prior authorization"], required_phrase_weights=[0.93, 0.79],
                      required_phrase_tag_codes=["j:payor.sunshine_health", "d:utilization_management.prior_authorization"]),
             bm25=dict(n_extra=30), expansion=dict(domain_tags=["d:utilization_management.prior_authorization"])),
        dict(req=dict(query="timely filing deadline for corrected claim", k=6, mode="corpus", required_phrases=["timely filing"],
                      required_phrase_weights=[0.8], required_phrase_tag_codes=["d:claims.timely_filing"], assembly_strategy="balanced",
                      canonical_floor=0.5), bm25=dict(n_extra=25)),
        dict(req=dict(query="provider services phone number", k=10, mode="corpus", assembly_strategy="canonical_first",
                      neighbor_paragraph_window=1, neighbor_page_window=0), bm25=dict(n_extra=20)),
        dict(req=dict(query="medical records documentation required", k=10, mode="precision", required_phrases=["medical records"],
                      required_phrase_weights=[0.7], required_phrase_tag_codes=[None]), bm25=dict(n_extra=40)),
        dict(req=dict(query="appeals", k=10, mode="precision", required_phrase_tag_codes=["d:claims.appeals"]), bm25=dict(n_extra=15)),
        dict(req=dict(query="behavioral health credentialing", k=200, mode="corpus", include_document_ids=[d["document_id"] for d in docs[:25]]),
             bm25=dict(n_extra=10)),
        dict(req=dict(query="anything", k=10, mode="recall", include_document_ids=[d["document_id"] for d in docs])),      # > 2000 chunks? no: 900
        dict(req=dict(query="   ", k=10)),
    ]
    out = []
    for ci, case in enumerate(cases):
        q = (X[int(rng.integers(0, len(rows)))].astype(np.float64) + 0.1 * rng.standard_normal(dim)).tolist() if ci % 2 == 0 \
            else rng.standard_normal(dim).tolist()
        qv = np.asarray([np.float32(x) for x in q], dtype=np.float32)
        order = [i for i in np.argsort(cd(X, qv), kind="stable") if rows[i]["has_vec"]][:6]
        bm = bm25_list(order[1:4], case["bm25"]["n_extra"]) if case.get("bm25") else []
        expansion = dict(matched_codes=[], expansion_phrases=[], expansion_phrases_count=0, final_tsquery="", log=[], domain_tags=[],
                         jurisdiction_tags=[], process_tags=[])
        expansion.update(case.get("expansion") or {})

        async def fake_bm25(db, query, k, filters, include_document_ids, search_id="", tag_mode="auto", _bm=bm, _e=expansion):
            return [dict(c) for c in _bm], None, dict(_e)

        async def fake_embed(query, search_id="", _q=q):
            return list(_q), 0.0, False

        async def fake_expand(db, query, _c=case):
            return ref_lex.LexiconExpansion(**_c["lexicon"]) if _c.get("lexicon") else None
        ref_cs._bm25_arm, ref_cs._embed_with_cache = fake_bm25, fake_embed
        ref_lex.expand_query_via_lexicon = fake_expand
        req = ref_cs.CorpusSearchRequest(**case["req"])

        async def run(req=req):
            db = Session()
            resp = await ref_cs.corpus_search(db, req)
            await asyncio.sleep(0)                      # let the fire-and-forget task run
            return resp, db.log
        resp, log = asyncio.run(run())
        tel = resp.telemetry
        out.append({"request": case["req"], "query_embedding": q, "bm25": bm, "bm25_expansion": expansion, "lexicon": case.get("lexicon"),
                    "statements": log, "chunks": [c.model_dump() for c in resp.chunks],
                    "telemetry": {k: tel.get(k) for k in ("mode", "k", "arm_hits", "candidates", "returned", "min_label_applied", "assembly", "error")}})
    dump_json(out, "corpus_search.json")
    print("corpus_search cases:", [(c["request"].get("mode", "corpus"), len(c["chunks"]), c["telemetry"].get("candidates")) for c in out])


def main():
    if not os.path.isdir(REF):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    install_stubs()
    sys.path.insert(0, REF)
    if sys.argv[1:] == ["pool"]:                       # regenerate one fixture without touching the others
        make_pool_golden()
        return
    from oracle import oracle
    if sys.argv[1:] == ["corpus_search"]:
        make_corpus_search_golden(oracle)
        return
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

    make_rerank_golden(oracle, ref_cs)
    make_pool_golden()
    make_corpus_search_golden(oracle)

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
    dump_json(table_json, "table.json")
    dump_json(arm_out, "vector_arm.json")
    dump_json(store_out, "store_search.json")
    print(f"wrote {len(arm_out)} _vector_arm cases, {len(store_out)} PgVectorStore cases; "
          f"result sizes {[len(c['result']) for c in arm_out]}")
    print("statements per case:", [len(c["statements"]) for c in arm_out])


if __name__ == "__main__":
    main()
