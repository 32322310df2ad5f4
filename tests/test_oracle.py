"""CPU tests of the oracle itself: the pgvector restatement against pgvector's published
known answers, C vs numpy, SQL ordering rules, and the statement-level restatements.

PARITY UNPINNED (see oracle/oracle.py): the reference holds no golden vectors for this path and
pgvector is not vendored; the known answers below are pgvector's own regression expectations
for cosine_distance (test/expected/functions.out, v0.5.1), restated from its published suite.
"""
import math

import numpy as np
import pytest

from mrag_b200 import synth

# (a, b, expected cosine_distance) -- pgvector test/sql/functions.sql, cosine_distance block
PGVECTOR_KAT = [
    ([1, 2], [2, 4], 0.0),
    ([1, 2], [0, 0], math.nan),
    ([1, 1], [1, 1], 0.0),
    ([1, 0], [0, 2], 1.0),
    ([1, 1], [-1, -1], 2.0),
    ([1, 1], [1.1, 1.1], 0.0),
    ([1, 1], [-1.1, -1.1], 2.0),
    ([3e38], [3e38], math.nan),
]


@pytest.mark.parametrize("a,b,want", PGVECTOR_KAT)
def test_pgvector_known_answers(oracle, a, b, want):
    A = np.asarray([a], dtype=np.float32)
    q = np.asarray(b, dtype=np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        got_c = oracle.cosine_distance_c(A, q)[0]
        got_np = oracle.cosine_distance_np(A, q)[0]
    for got in (got_c, got_np):
        if math.isnan(want):
            assert math.isnan(got)
        else:
            assert got == pytest.approx(want, abs=1e-7)


def test_c_matches_numpy(oracle):
    X, _ = synth.make_corpus(3000, 96, seed=5)
    Q = synth.make_queries(X, 4, seed=6)
    for q in Q:
        dc = oracle.cosine_distance_c(X, q)
        dn = oracle.cosine_distance_np(X, q)
        both = ~(np.isnan(dc) | np.isnan(dn))
        assert (np.isnan(dc) == np.isnan(dn)).all()
        # float32 accumulation in unspecified order: a few ulp of fp32 on a sum of 96 terms
        assert np.abs(dc[both] - dn[both]).max() < 5e-6


def test_order_by_nan_last_ties_by_row(oracle):
    dist = np.array([0.5, np.nan, 0.25, 0.5, 0.25, np.nan, 1.5])
    rows, sims = oracle.order_by_limit(dist, None, 7)
    assert rows.tolist() == [2, 4, 0, 3, 6, 1, 5]
    assert np.isnan(sims[-2:]).all() and sims[0] == 0.75
    rows, _ = oracle.order_by_limit(dist, np.array([1, 1, 0, 1, 0, 0, 1], bool), 3)
    assert rows.tolist() == [0, 3, 6]
    # the C top-N heap gives the same answer
    r = np.full(7, -1, np.int64); s = np.full(7, np.nan)
    import ctypes
    m = oracle.clib().pgv_topk(dist.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), None, 7, 7,
                               r.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                               s.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert m == 7 and r.tolist() == [2, 4, 0, 3, 6, 1, 5]


def test_search_c_vs_numpy_paths(oracle):
    X, valid = synth.make_corpus(4000, 64, seed=11, null_frac=5e-3)
    Q = synth.make_queries(X, 6, seed=12)
    mask = valid.astype(bool)
    r1, s1, c1 = oracle.search(X, Q, 25, mask, use_c=True)
    for i in range(Q.shape[0]):
        sim_all = oracle.all_similarities(X, Q[i])
        oracle.check_topk(r1[i], s1[i], int(c1[i]), sim_all, mask, 25, rtol=1e-6)
    r2, s2, c2 = oracle.search(X, Q, 25, mask, use_c=False)
    assert (c1 == c2).all()
    for i in range(Q.shape[0]):
        sim_all = oracle.all_similarities(X, Q[i])
        oracle.check_topk(r2[i], s2[i], int(c2[i]), sim_all, mask, 25, rtol=1e-5, tie_tol=2e-6)


def test_round_bf16_matches_torch(oracle):
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37.0
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert (oracle.round_bf16(x) == want).all()
    import ctypes
    out = np.empty_like(x)
    oracle.clib().pgv_round_bf16(x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), x.size)
    assert (out == want).all()


def test_to_float4_is_strtof(oracle):
    # repr(float(x)) -> strtof == nearest float32 of the python float
    vals = [0.1, 1 / 3, 1e-40, 123456.789, -2.5e-7]
    got = oracle.to_float4(vals)
    assert got.dtype == np.float32
    assert [float(np.float32(float(repr(v)))) for v in vals] == got.astype(np.float64).tolist()


def test_ilike(oracle):
    assert oracle.sql_ilike("Molina Healthcare of Florida", "%molina healthcare%")
    assert oracle.sql_ilike("AHCA", "%ahca%")
    assert not oracle.sql_ilike("Humana", "%ahca%")
    assert oracle.sql_ilike("a_b", "a_b") and oracle.sql_ilike("axb", "a_b")
    assert oracle.sql_ilike("", "%%") and not oracle.sql_ilike("", "%x%")


def test_vector_arm_restatement_semantics(oracle):
    """Hand-checkable table: LIMIT, strict->relaxed retry, clamp, min_similarity, stop-at-k."""
    from helpers import build_tables
    from mrag_b200.corpus_search import CorpusFilters, LexiconExpansion
    ot, _, X, valid, meta, info = build_tables(oracle, 600, 32, seed=3, with_product=False)
    q = (X[10] * 1.0).tolist()
    res = oracle.vector_arm(ot, q, 5, None, None)
    assert len(res) == 5 and res[0]["id"] == ot.id[10] or res[0]["similarity"] == pytest.approx(1.0, abs=1e-6)
    assert all(0.0 <= r["similarity"] <= 1.0 and r["_arm"] == "vector" for r in res)
    sims = [r["similarity"] for r in res]
    assert sims == sorted(sims, reverse=True)
    # min_similarity drops everything but the near-duplicates
    res2 = oracle.vector_arm(ot, q, 5, None, None, min_similarity=0.99, over_fetch_factor=8)
    assert 1 <= len(res2) <= 5 and all(r["similarity"] >= 0.99 for r in res2)
    # a payer filter only returns that payer (+ the FL union for MCO payers)
    res3 = oracle.vector_arm(ot, q, 50, CorpusFilters(payer="Sunshine Health"), None)
    assert res3 and all(r["payer"] == "Sunshine Health" or (r["payer"] in oracle.FL_STATE_AUTHORITY_PAYERS and r["state"] == "FL")
                        for r in res3)
    # strict filter that matches nothing falls back to relaxed in auto mode, not in strict mode
    exp = LexiconExpansion(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"])
    auto = oracle.vector_arm(ot, q, 5, None, None, expansion=exp, tag_mode="auto")
    strict = oracle.vector_arm(ot, q, 5, None, None, expansion=exp, tag_mode="strict")
    assert strict == [] and len(auto) > 0
    docs_with_tag = {d for d, ks in ot.doc_d_tags.items() if "topic_000.leaf" in ks}
    assert all(r["document_id"] in docs_with_tag for r in auto)
