"""GPU parity tests: the CUDA path (through the C ABI of include/mrag.h) against the CPU oracle on
the same seeded inputs.  Bar (BASELINE.json north_star): returned ids and their order identical
except inside groups of tied scores; scores within 1e-4 relative for the fp32 corpus, 1e-2 for the
bf16 corpus (where the oracle sees the bf16-rounded corpus upcast to fp32)."""
import math
import threading

import numpy as np
import pytest

import mrag_b200
from mrag_b200 import _native as N
from mrag_b200 import synth
from mrag_b200.index import Filter, Index, make_meta, merge_topk

pytestmark = pytest.mark.gpu

RTOL = {"f32": 1e-4, "bf16": 1e-2}
# in bf16 mode the scores are still computed with fp32 accumulation from the rounded rows, so ties
# are judged as tightly as in fp32 mode
TIE_TOL = 1e-6


def check_all(oracle, Xs, Q, mask, k, scores, rows, counts, dtype, row_base=0):
    for i in range(Q.shape[0]):
        sim_all = oracle.all_similarities(Xs, Q[i])
        r = rows[i].copy()
        r[r >= 0] -= row_base
        oracle.check_topk(r, scores[i], int(counts[i]), sim_all, mask, k, rtol=RTOL[dtype], tie_tol=TIE_TOL)


def stored(oracle, X, dtype):
    return oracle.round_bf16(X) if dtype == "bf16" else X


# ---------------------------------------------------------------------------------------------
# C1: 100k x 768, single query, top-10 (BASELINE.json configs[0])
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_c1_100k_768_top10(oracle, dtype):
    n, dim, k = 100_000, 768, 10
    X, valid = synth.make_corpus(n, dim, seed=1234)
    Q = synth.make_queries(X, 6, seed=4321)
    idx = Index(dim, dtype, 0, n)
    idx.append(X, make_meta(n, valid=valid))
    Xs = stored(oracle, X, dtype)
    for i in range(Q.shape[0]):                       # single-query calls, as /api/query issues them
        s, r, c = idx.search(Q[i:i + 1], k)
        check_all(oracle, Xs, Q[i:i + 1], valid.astype(bool), k, s, r, c, dtype)
    assert idx.last_scan_kind() in ("gemv", "mma")
    assert idx.last_kernel_ms(1) > 0
    idx.close()


# ---------------------------------------------------------------------------------------------
# shapes: ragged dims, batch sizes, k (incl. the multi-round path k > MRAG_FUSED_K)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,dim,nq,k", [
    (5000, 768, 1, 10), (5000, 1536, 3, 10), (4097, 100, 5, 7), (33, 1, 2, 5), (31, 3, 1, 40),
    (20000, 64, 64, 10), (20000, 768, 7, 100), (9000, 128, 4, 128), (9000, 128, 2, 129),
    (6000, 96, 3, 300), (12000, 64, 1, 1600), (2000, 768, 17, 1),
])
def test_shapes(oracle, dtype, n, dim, nq, k):
    X, valid = synth.make_corpus(n, dim, seed=n + dim, null_frac=3e-3)
    Q = synth.make_queries(X, nq, seed=k)
    idx = Index(dim, dtype, 0, n + 5)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k)
    check_all(oracle, stored(oracle, X, dtype), Q, valid.astype(bool), k, s, r, c, dtype)
    idx.close()


def test_empty_and_tiny_index(oracle):
    idx = Index(16, "f32", 0, 100)
    s, r, c = idx.search(np.ones((2, 16), np.float32), 5)
    assert (c == 0).all() and (r == -1).all() and np.isnan(s).all()
    X = np.eye(16, dtype=np.float32)[:3]
    idx.append(X)
    s, r, c = idx.search(X[1:2] * 7.0, 5)
    assert c[0] == 3 and r[0, 0] == 1 and s[0, 0] == pytest.approx(1.0, abs=1e-6)
    assert sorted(r[0, 1:3].tolist()) == [0, 2] and r[0, 1] == 0 and (r[0, 3:] == -1).all()
    idx.close()


def test_incremental_append_matches_one_shot(oracle):
    n, dim = 7001, 200
    X, valid = synth.make_corpus(n, dim, seed=77)
    Q = synth.make_queries(X, 4, seed=78)
    idx = Index(dim, "f32", 0, n)
    cuts = [0, 50, 51, 3000, 3033, n]          # the worker appends batches of 50 (embedding_worker.py:229-266)
    for lo, hi in zip(cuts, cuts[1:]):
        first = idx.append(X[lo:hi], make_meta(hi - lo, doc_idx=np.arange(lo, hi), valid=valid[lo:hi]))
        assert first == lo
    assert len(idx) == n
    s, r, c = idx.search(Q, 20)
    check_all(oracle, X, Q, valid.astype(bool), 20, s, r, c, "f32")
    with pytest.raises(N.MragError) as e:
        idx.append(X[:1])
    assert e.value.code == N.MRAG_ERR_OOM
    idx.close()


# ---------------------------------------------------------------------------------------------
# NaN / NULL / ties
# ---------------------------------------------------------------------------------------------
def test_nan_rows_sort_last_and_null_rows_never_return(oracle):
    n, dim, k = 300, 24, 300
    X, valid = synth.make_corpus(n, dim, seed=5, null_frac=0.05, zero_norm_rows=4)
    Q = synth.make_queries(X, 3, seed=6)
    idx = Index(dim, "f32", 0, n)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k)
    mask = valid.astype(bool)
    check_all(oracle, X, Q, mask, k, s, r, c, "f32")
    zero = np.nonzero((np.abs(X).sum(axis=1) == 0) & mask)[0]
    assert len(zero) >= 1
    for i in range(3):
        assert c[i] == mask.sum()
        tail = r[i, c[i] - len(zero):c[i]]
        assert tail.tolist() == zero.tolist() and np.isnan(s[i, c[i] - len(zero):c[i]]).all()
        assert not set(r[i, :c[i]].tolist()) & set(np.nonzero(~mask)[0].tolist())
    # a zero query: every distance is NaN -> the first k passing rows in row order
    s, r, c = idx.search(np.zeros((1, dim), np.float32), 10)
    assert c[0] == 10 and r[0].tolist() == np.nonzero(mask)[0][:10].tolist() and np.isnan(s[0]).all()
    idx.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_duplicate_cluster_ties_break_by_row(oracle, dtype):
    n, dim, k = 6000, 256, 50
    X, valid = synth.make_corpus(n, dim, seed=9, dup_frac=0.0)
    rng = np.random.default_rng(1)
    dup_rows = np.sort(rng.choice(n, size=120, replace=False))
    X[dup_rows] = X[dup_rows[0]]
    valid[dup_rows] = 1
    q = X[dup_rows[0]][None, :] * 3.0
    idx = Index(dim, dtype, 0, n)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(q, k)
    assert r[0].tolist() == dup_rows[:k].tolist()          # all tie at 1.0: ascending row
    assert np.allclose(s[0], 1.0, atol=1e-6)
    check_all(oracle, stored(oracle, X, dtype), q, valid.astype(bool), k, s, r, c, dtype)
    idx.close()


def test_scores_clamped_to_unit_range():
    X = np.full((40, 7), 0.1, dtype=np.float32)
    X[1::2] *= -1
    idx = Index(7, "f32", 0, 40)
    idx.append(X)
    s, r, c = idx.search(X[:1], 40)
    assert s.max() <= 1.0 and s.min() >= -1.0 and c[0] == 40
    assert r[0, :20].tolist() == list(range(0, 40, 2)) and r[0, 20:].tolist() == list(range(1, 40, 2))
    idx.close()


# ---------------------------------------------------------------------------------------------
# WHERE clauses on codes (K2) against numpy on the same codes
# ---------------------------------------------------------------------------------------------
def test_filter_clauses_codes(oracle):
    n, dim, k = 30000, 64, 25
    X, valid = synth.make_corpus(n, dim, seed=31, null_frac=2e-3)
    meta, doc_tags, info = synth.make_metadata(n, seed=32, rows_per_doc=64, valid=valid)
    Q = synth.make_queries(X, 3, seed=33)
    idx = Index(dim, "f32", 0, n)
    idx.append(X, meta)
    idx.set_doc_tags(0, doc_tags)
    v = valid.astype(bool)
    P, S = synth.PAYERS, synth.STATES
    fl = S.index("FL")
    alt = [P.index(p) for p in ("AHCA", "Ahca.myflorida", "Florida Medicaid")]
    doc = meta["doc_idx"]
    rng = np.random.default_rng(3)
    pools = [rng.choice(info["n_docs"], size=m, replace=False) for m in (1, 19, 50, 300)]
    tagbits = lambda bits: np.array([any(info["tagmat"][d, b] for b in bits) for d in range(info["n_docs"])])[doc]
    cases = [
        ("payer", Filter().payer_in([P.index("Humana")]), meta["payer"] == P.index("Humana")),
        ("payer+FL union", Filter().payer_in([P.index("Sunshine Health")], alt, fl),
         (meta["payer"] == P.index("Sunshine Health")) | (np.isin(meta["payer"], alt) & (meta["state"] == fl))),
        ("state", Filter().state_eq(S.index("TX")), meta["state"] == S.index("TX")),
        ("program", Filter().program_eq(3), meta["program"] == 3),
        ("authority", Filter().authority_eq(1), meta["authority"] == 1),
        ("source_type", Filter().source_type_eq(2), meta["source_type"] == 2),
        ("unknown value", Filter().payer_in([0xFFFE]), np.zeros(n, bool)),
        ("doc_eq", Filter().doc_eq(int(doc[n // 2])), doc == doc[n // 2]),
        ("tag relaxed 10%", Filter().tag_relaxed([0, 3]), tagbits([0, 3])),
        ("tag relaxed 0.1%", Filter().tag_relaxed([2]), tagbits([2])),
        ("tag strict", Filter().tag_strict([S.index("GA")], [5], [P.index("Aetna")]),
         (meta["state"] == S.index("GA")) | (meta["program"] == 5) | (meta["payer"] == P.index("Aetna"))),
        ("combo", Filter().payer_in([P.index("Sunshine Health")], alt, fl).authority_eq(0).tag_relaxed([0, 1, 3, 6]),
         ((meta["payer"] == P.index("Sunshine Health")) | (np.isin(meta["payer"], alt) & (meta["state"] == fl)))
         & (meta["authority"] == 0) & tagbits([0, 1, 3, 6])),
        ("empty pool", Filter().doc_pool([]), np.zeros(n, bool)),
    ] + [(f"pool{len(p)}", Filter().doc_pool(p), np.isin(doc, p)) for p in pools]
    for name, flt, want in cases:
        want = want & v
        bits, n_pass = idx.filter_mask(flt)
        got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (got == want).all(), name
        assert n_pass == int(want.sum()), name
        s, r, c = idx.search(Q, k, flt)
        check_all(oracle, X, Q, want, k, s, r, c, "f32")
    bits, n_pass = idx.filter_mask(None)
    assert n_pass == int(v.sum())
    idx.close()


def test_tombstone_document(oracle):
    n, dim = 5000, 48
    X, valid = synth.make_corpus(n, dim, seed=41)
    meta, _, info = synth.make_metadata(n, seed=42, rows_per_doc=32, valid=valid)
    Q = synth.make_queries(X, 2, seed=43)
    idx = Index(dim, "f32", 0, n)
    idx.append(X, meta)
    d = int(meta["doc_idx"][2500])
    want_gone = (meta["doc_idx"] == d) & valid.astype(bool)
    assert idx.live_rows() == (n, int(valid.sum()))
    assert idx.tombstone_doc(d) == int(want_gone.sum())
    assert idx.tombstone_doc(d) == 0
    assert idx.live_rows() == (n - int((meta["doc_idx"] == d).sum()), int(valid.sum()) - int(want_gone.sum()))
    mask = valid.astype(bool) & ~want_gone
    s, r, c = idx.search(Q, 30)
    check_all(oracle, X, Q, mask, 30, s, r, c, "f32")
    idx.close()


# ---------------------------------------------------------------------------------------------
# the plugin boundary: B200VectorStore and vector_arm against the statement-level oracle
# ---------------------------------------------------------------------------------------------
def _same_hits(got, want, key, rtol):
    assert [g["id"] for g in got] == [w["id"] for w in want]
    for g, w in zip(got, want):
        assert {x: g[x] for x in g if x != key} == {x: w[x] for x in w if x != key}
        if isinstance(w[key], float) and math.isnan(w[key]):
            assert math.isnan(g[key])
        else:
            assert g[key] == pytest.approx(w[key], rel=rtol, abs=rtol * 1e-3)


@pytest.fixture(scope="module")
def tables(oracle):
    from helpers import build_tables
    ot, pt, X, valid, meta, info = build_tables(oracle, 6000, 96, seed=7, dtype="f32")
    return ot, pt, X, valid


def test_store_search_matches_pg_statement(oracle, tables):
    ot, pt, X, valid = tables
    store = mrag_b200.B200VectorStore(table=pt)
    rng = np.random.default_rng(0)
    doc = ot.document_id[1234]
    cases = [
        dict(k=10), dict(k=1), dict(k=100), dict(k=10, document_id=doc), dict(k=5, document_id="not-a-doc"),
        dict(k=10, filters={"payer": "Humana"}), dict(k=10, filters={"state": "FL", "authority_level": "payer_policy"}),
        dict(k=10, filters={"payer": "", "state": None, "bogus": "x"}),           # all skipped (vector_store.py:250-258)
        dict(k=10, filters={"source_type": "fact", "payer": "Nobody"}),
        dict(k=10, filters={"document_id": doc}), dict(k=10, document_id=doc, filters={"document_id": ot.document_id[0]}),
    ]
    for i, kw in enumerate(cases):
        emb = (X[int(rng.integers(0, len(X)))] + 0.05 * rng.standard_normal(X.shape[1])).tolist() if i % 2 else \
            rng.standard_normal(X.shape[1]).tolist()
        want = oracle.pg_store_search(ot, emb, kw["k"], kw.get("document_id"), kw.get("filters"))
        got = store.search(emb, kw["k"], kw.get("document_id"), kw.get("filters"))
        _same_hits(got, want, "distance", 1e-4)
        assert all(set(g) == {"id", "document_id", "source_type", "source_id", "distance"} for g in got)
    import asyncio
    emb = X[77].tolist()
    assert [g["id"] for g in asyncio.run(store.asearch(emb, 10))] == [w["id"] for w in oracle.pg_store_search(ot, emb, 10)]
    with pytest.raises(ValueError):
        store.search(emb[:-1], 10)


def test_vector_arm_matches_reference_restatement(oracle, tables):
    from mrag_b200.corpus_search import CorpusFilters, LexiconExpansion
    ot, pt, X, valid = tables
    rng = np.random.default_rng(1)
    docs = sorted(set(ot.document_id))
    pool19 = [docs[i] for i in rng.choice(len(docs), 19, replace=False)]
    pool_big = [docs[i] for i in rng.choice(len(docs), min(300, len(docs)), replace=False)] + ["ffffffff-0000-0000-0000-000000000000"]
    E = LexiconExpansion
    cases = [
        dict(k=10), dict(k=20, over_fetch_factor=8, min_similarity=0.3), dict(k=200, over_fetch_factor=8),
        dict(k=10, filters=CorpusFilters(payer="Sunshine Health")), dict(k=10, filters=CorpusFilters(payer="Centene", state="FL")),
        dict(k=10, filters=CorpusFilters(program="Medicaid", authority_level="payer_policy")),
        dict(k=10, include_document_ids=pool19), dict(k=40, include_document_ids=pool_big, filters=CorpusFilters(state="FL")),
        dict(k=10, expansion=E(jurisdiction_tags=["j:state.fl", "j:bad"]), tag_mode="auto"),
        dict(k=10, expansion=E(jurisdiction_tags=["j:payor.molina_healthcare", "j:program.medic"]), tag_mode="strict"),
        dict(k=10, expansion=E(jurisdiction_tags=["j:regulatory_authority.ahca"], domain_tags=["d:topic_000.leaf"]), tag_mode="auto"),
        dict(k=10, expansion=E(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"], process_tags=["p:topic_003.leaf"]), tag_mode="auto"),
        dict(k=10, expansion=E(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"]), tag_mode="strict"),
        dict(k=10, expansion=E(domain_tags=["d:topic_006.leaf", "d:unknown.key"], process_tags=["p:topic_001.leaf"]), tag_mode="relaxed"),
        dict(k=10, expansion=E(jurisdiction_tags=["j:state.fl"], domain_tags=["d:topic_000.leaf"]), tag_mode="none"),
        dict(k=10, expansion=E(domain_tags=["d:topic_000.leaf"]), tag_mode="auto"),      # strict empty -> unfiltered, no retry
        dict(k=5, min_similarity=0.999),
    ]
    for i, kw in enumerate(cases):
        emb = (X[int(rng.integers(0, len(X)))] + 0.05 * rng.standard_normal(X.shape[1])).tolist() if i % 2 == 0 else \
            rng.standard_normal(X.shape[1]).tolist()
        args = (emb, kw["k"], kw.get("filters"), kw.get("include_document_ids"))
        opt = dict(expansion=kw.get("expansion"), tag_mode=kw.get("tag_mode", "auto"),
                   min_similarity=kw.get("min_similarity"), over_fetch_factor=kw.get("over_fetch_factor", 1))
        want = oracle.vector_arm(ot, *args, **opt)
        got = mrag_b200.vector_arm(pt, *args, search_id=f"case{i}", **opt)
        assert [g["id"] for g in got] == [w["id"] for w in want], f"case {i}"
        for g, w in zip(got, want):
            assert set(g) == set(w)
            for key in w:
                if key in ("similarity", "match_score"):
                    assert g[key] == pytest.approx(w[key], rel=1e-4, abs=1e-7)
                else:
                    assert g[key] == w[key], (i, key)
    # NaN quirk of corpus_search.py:1569: a NaN similarity reports 1.0
    z = int(np.nonzero((np.abs(X).sum(axis=1) == 0) & valid.astype(bool))[0][0])
    # (NaN distances sort last, so ask for more rows than the document has)
    want = oracle.vector_arm(ot, X[5].tolist(), 100, None, [ot.document_id[z]])
    got = mrag_b200.vector_arm(pt, X[5].tolist(), 100, None, [ot.document_id[z]])
    assert [g["id"] for g in got] == [w["id"] for w in want]
    assert any(g["id"] == ot.id[z] and g["similarity"] == 1.0 for g in got)
    # over the LIMIT cap -> logged and [] (fail-soft), never raised
    assert mrag_b200.vector_arm(pt, X[5].tolist(), 300, None, None, over_fetch_factor=8) == []


def test_republish_document(oracle):
    """DELETE + INSERT of one document (publish.py:310-313): old rows vanish, new rows are found."""
    from helpers import build_tables
    ot, pt, X, valid, meta, info = build_tables(oracle, 800, 32, seed=13)
    store = mrag_b200.B200VectorStore(table=pt)
    doc = ot.document_id[400]
    old_ids = {ot.id[i] for i in range(800) if ot.document_id[i] == doc}
    store.delete_by_document(doc)
    emb = X[400].tolist()
    assert not {h["id"] for h in store.search(emb, 50)} & old_ids
    store.add(["new-1"], [emb], [{"document_id": doc, "source_type": "fact", "source_id": "s"}])
    hits = store.search(emb, 3)
    assert hits[0]["id"] == "new-1" and hits[0]["distance"] == pytest.approx(1.0, abs=1e-6)
    assert [h["id"] for h in store.search(emb, 5, document_id=doc)] == ["new-1"]


# ---------------------------------------------------------------------------------------------
# device I/O, re-entrancy, K4
# ---------------------------------------------------------------------------------------------
def test_device_io_and_row_base(oracle):
    import torch
    n, dim, k = 8000, 128, 16
    X, valid = synth.make_corpus(n, dim, seed=51)
    Q = synth.make_queries(X, 9, seed=52)
    idx = Index(dim, "bf16", 0, n)
    idx.append_device(torch.from_numpy(X).cuda(), make_meta(n, valid=valid))
    idx.set_row_base(1_000_000)
    qd = torch.from_numpy(Q).cuda()
    s, r, c = idx.search_device(qd, k)
    s2, r2, c2 = idx.search_device(qd, k, sync=False)
    torch.cuda.synchronize()
    assert torch.equal(r, r2) and torch.equal(c, c2) and torch.equal(s, s2)
    check_all(oracle, oracle.round_bf16(X), Q, valid.astype(bool), k, s.cpu().numpy(), r.cpu().numpy(),
              c.cpu().numpy(), "bf16", row_base=1_000_000)
    idx.close()


def test_concurrent_searches_on_one_handle(oracle):
    """Up to 5 narrow searches run concurrently per request (corpus_search_agent.py:794-797)."""
    n, dim, k = 20000, 96, 10
    X, valid = synth.make_corpus(n, dim, seed=61)
    Q = synth.make_queries(X, 40, seed=62)
    idx = Index(dim, "f32", 0, n)
    idx.append(X, make_meta(n, valid=valid))
    want_s, want_r, want_c = idx.search(Q, k)
    errs = []

    def work(t):
        try:
            for rep in range(5):
                for i in range(t, 40, 8):
                    s, r, c = idx.search(Q[i:i + 1], k)
                    assert (r[0] == want_r[i]).all() and c[0] == want_c[i]
        except Exception as e:   # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    check_all(oracle, X, Q[:5], valid.astype(bool), k, want_s[:5], want_r[:5], want_c[:5], "f32")
    idx.close()


def test_merge_topk_kernel():
    import torch
    rng = np.random.default_rng(5)
    for n_lists, nq, k in [(2, 3, 10), (8, 5, 100), (3, 1, 1), (8, 2, 1600)]:
        stride = nq * k + 13
        sc = np.full((n_lists, stride), np.nan, np.float32)
        ro = np.full((n_lists, stride), -1, np.int64)
        co = np.zeros((n_lists, nq + 3), np.int32)
        want = []
        for q in range(nq):
            cand = []
            for l in range(n_lists):
                m = int(rng.integers(0, k + 1))
                vals = np.sort(rng.choice(np.round(rng.standard_normal(40), 1), size=m).astype(np.float32))[::-1]
                n_nan = int(rng.integers(0, 3)) if m > 2 else 0
                vals = vals.copy()
                if n_nan:
                    vals[m - n_nan:] = np.nan
                rws = rng.choice(10_000_000, size=m, replace=False).astype(np.int64) + l * 10_000_000
                # lists arrive sorted: score desc, NaN last, row asc inside ties
                order = sorted(range(m), key=lambda j: (math.isnan(vals[j]), -vals[j] if not math.isnan(vals[j]) else 0, rws[j]))
                vals, rws = vals[order], rws[order]
                sc[l, q * k:q * k + m] = vals
                ro[l, q * k:q * k + m] = rws
                co[l, q] = m
                cand += list(zip(vals.tolist(), rws.tolist()))
            cand.sort(key=lambda t: (math.isnan(t[0]), -t[0] if not math.isnan(t[0]) else 0, t[1]))
            want.append(cand[:k])
        s, r, c = merge_topk(0, torch.from_numpy(sc).cuda(), torch.from_numpy(ro).cuda(), torch.from_numpy(co).cuda(),
                             n_lists, nq, k, (stride, stride, nq + 3))
        torch.cuda.synchronize()
        s, r, c = s.cpu().numpy(), r.cpu().numpy(), c.cpu().numpy()
        for q in range(nq):
            assert c[q] == len(want[q])
            assert r[q, :c[q]].tolist() == [w[1] for w in want[q]]
            got_s = s[q, :c[q]]
            ws = np.array([w[0] for w in want[q]], np.float32)
            assert ((got_s == ws) | (np.isnan(got_s) & np.isnan(ws))).all()
            assert (r[q, c[q]:] == -1).all()


def test_bad_arguments():
    idx = Index(8, "f32", 0, 10)
    with pytest.raises(N.MragError):
        idx.search(np.zeros((1, 8), np.float32), 0)
    with pytest.raises(N.MragError):
        idx.search(np.zeros((1, 8), np.float32), N.MRAG_MAX_K + 1)
    with pytest.raises(ValueError):
        idx.search(np.zeros((1, 9), np.float32), 1)
    with pytest.raises(ValueError):
        idx.append(np.full((1, 8), np.nan, np.float32))
    with pytest.raises(N.MragError):
        Index(8, "f32", 99, 10)
    idx.close()


# ---------------------------------------------------------------------------------------------
# the tensor-core scan (TMA + tcgen05 + TMEM), pinned with MRAG_OPT_FORCE_MMA, and the CUDA-core
# scan pinned with MRAG_OPT_FORCE_GEMV, must both equal the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dim,nq,k", [
    (5000, 768, 1, 10), (5000, 768, 2, 10), (64, 64, 3, 10), (63, 64, 1, 5), (65, 128, 64, 10),
    (20000, 768, 64, 10), (20000, 768, 65, 10), (30000, 256, 130, 10), (20000, 768, 7, 100),
    (9000, 128, 4, 128), (9000, 128, 2, 129), (6000, 96, 3, 300), (4097, 100, 5, 7), (12000, 64, 33, 64),
    (100000, 768, 16, 10),
])
@pytest.mark.parametrize("path", ["mma", "gemv"])
def test_scan_paths_bf16(oracle, path, n, dim, nq, k):
    X, valid = synth.make_corpus(n, dim, seed=n + dim + 1, null_frac=3e-3)
    Q = synth.make_queries(X, nq, seed=k + 1)
    idx = Index(dim, "bf16", 0, n + 5)
    idx.append(X, make_meta(n, valid=valid))
    opt = N.OPT_FORCE_MMA if path == "mma" else N.OPT_FORCE_GEMV
    s, r, c = idx.search(Q, k, options=opt)
    assert idx.last_scan_kind() == path
    check_all(oracle, oracle.round_bf16(X), Q, valid.astype(bool), k, s, r, c, "bf16")
    idx.close()


def test_mma_default_dispatch_and_limits(oracle):
    X, valid = synth.make_corpus(3000, 768, seed=3)
    Q = synth.make_queries(X, 4, seed=4)
    idx = Index(768, "bf16", 0, 3000)
    idx.append(X, make_meta(3000, valid=valid))
    idx.search(Q, 10)
    assert idx.last_scan_kind() == "mma"          # batches go to the tensor cores
    idx.search(Q[:1], 10)
    assert idx.last_scan_kind() == "mma"          # an unfiltered single query too (it streams faster there)
    idx.search(Q[:1], 10, Filter().doc_eq(3))
    assert idx.last_scan_kind() == "gemv"         # a filtered single query: the CUDA-core scan skips rows one by one
    idx.close()
    for dtype, dim in (("f32", 768), ("bf16", 1600)):       # outside the tensor-core scan's envelope
        idx = Index(dim, dtype, 0, 100)
        idx.append(np.ones((3, dim), np.float32))
        idx.search(np.ones((2, dim), np.float32), 2)
        assert idx.last_scan_kind() == "gemv"
        with pytest.raises(N.MragError):
            idx.search(np.ones((2, dim), np.float32), 2, options=N.OPT_FORCE_MMA)
        idx.close()
    # rows of 769 .. 1536 elements (the reference's production dimension is 1536): k-split CTA pairs
    idx = Index(1536, "bf16", 0, 100)
    idx.append(np.ones((3, 1536), np.float32))
    idx.search(np.ones((2, 1536), np.float32), 2)
    assert idx.last_scan_kind() == "mma_ks"
    idx.close()


# ---------------------------------------------------------------------------------------------
# rows of 769 .. 1536 elements: the k-split pair kernel (two CTAs share every tile along K, the helper's
# partial dots travel through distributed shared memory) must equal the oracle like the one-CTA kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dim,nq,k", [
    (5000, 1536, 1, 10), (5000, 1536, 2, 10), (64, 832, 3, 10), (63, 1024, 1, 5), (65, 1536, 64, 10),
    (20000, 1536, 64, 10), (20000, 1536, 65, 10), (30000, 1024, 130, 10), (20000, 832, 7, 100),
    (9000, 1472, 4, 128), (9000, 900, 2, 129), (6000, 1500, 3, 300), (4097, 770, 5, 7), (12000, 1280, 33, 64),
    (100000, 1536, 16, 10), (9473, 1536, 40, 10), (9472 * 2 + 1, 1152, 64, 16),
])
def test_ksplit_scan_bf16(oracle, n, dim, nq, k):
    X, valid = synth.make_corpus(n, dim, seed=n + dim + 1, null_frac=3e-3)
    Q = synth.make_queries(X, nq, seed=k + 1)
    idx = Index(dim, "bf16", 0, n + 5)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA)
    assert idx.last_scan_kind() == "mma_ks"
    check_all(oracle, oracle.round_bf16(X), Q, valid.astype(bool), k, s, r, c, "bf16")
    s2, r2, c2 = idx.search(Q, k)                        # default dispatch (single queries included) -> same bytes
    assert idx.last_scan_kind() == "mma_ks"
    assert (s2.tobytes(), r2.tobytes(), c2.tobytes()) == (s.tobytes(), r.tobytes(), c.tobytes())
    idx.close()


def test_ksplit_with_filters_ties_and_nan(oracle):
    n, dim, k = 40000, 1536, 20
    X, valid = synth.make_corpus(n, dim, seed=231, null_frac=2e-3, zero_norm_rows=3)
    dup = np.arange(1000, 1000 + 90)                  # a 90-row boilerplate cluster across two tiles (= both leaders)
    X[dup] = X[dup[0]]
    meta, doc_tags, info = synth.make_metadata(n, seed=232, rows_per_doc=64, valid=valid)
    Q = synth.make_queries(X, 9, seed=233)
    Q[0] = X[dup[0]] * 2.0
    Q[1] = 0.0                                        # zero query: all NaN
    idx = Index(dim, "bf16", 0, n)
    idx.append(X, meta)
    idx.set_doc_tags(0, doc_tags)
    Xs = oracle.round_bf16(X)
    v = valid.astype(bool)
    doc = meta["doc_idx"]
    pool = np.random.default_rng(5).choice(info["n_docs"], size=40, replace=False)     # most tiles are skipped entirely
    cases = [
        (None, v),
        (Filter().doc_pool(pool), np.isin(doc, pool) & v),
        (Filter().payer_in([2]), (meta["payer"] == 2) & v),
        (Filter().doc_eq(int(doc[dup[0]])), (doc == doc[dup[0]]) & v),
        (Filter().payer_in([0xFFFE]), np.zeros(n, bool)),
    ]
    for flt, want in cases:
        s, r, c = idx.search(Q, k, flt, options=N.OPT_FORCE_MMA)
        assert idx.last_scan_kind() == "mma_ks"
        check_all(oracle, Xs, Q, want, k, s, r, c, "bf16")
    idx.close()


def test_ksplit_threshold_sampling_pass(oracle):
    """the sampled admission bound + the shared-memory buffer select of the k-split kernel on a small shard"""
    import subprocess, sys, textwrap, os
    code = textwrap.dedent('''
        import numpy as np, sys
        sys.path.insert(0, %r)
        import mrag_b200
        from mrag_b200 import synth, _native as N
        from mrag_b200.index import Index, make_meta, Filter
        from oracle import oracle
        n, dim = 60000, 1536
        X, valid = synth.make_corpus(n, dim, seed=177, null_frac=1e-3)
        dup = np.arange(0, n, 64 * 64)[:10] + 3
        X[dup] = X[dup[0]]; valid[dup] = 1
        Q = synth.make_queries(X, 70, seed=178)
        Q[0] = X[dup[0]]
        idx = Index(dim, "bf16", 0, n)
        idx.append(X, make_meta(n, valid=valid))
        Xs = oracle.round_bf16(X)
        for k in (10, 100, 200):
            s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA)
            assert idx.last_scan_kind() == "mma_ks"
            for i in range(Q.shape[0]):
                oracle.check_topk(r[i], s[i], int(c[i]), oracle.all_similarities(Xs, Q[i]), valid.astype(bool), k, rtol=1e-2)
        print("SAMPLING-OK")
    ''') % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MRAG_SAMPLE_MIN_TILES="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "SAMPLING-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_mma_with_filters_ties_and_nan(oracle):
    n, dim, k = 40000, 768, 20
    X, valid = synth.make_corpus(n, dim, seed=131, null_frac=2e-3, zero_norm_rows=3)
    dup = np.arange(1000, 1000 + 90)                  # a 90-row boilerplate cluster across two tiles
    X[dup] = X[dup[0]]
    meta, doc_tags, info = synth.make_metadata(n, seed=132, rows_per_doc=64, valid=valid)
    Q = synth.make_queries(X, 9, seed=133)
    Q[0] = X[dup[0]] * 2.0
    Q[1] = 0.0                                        # zero query: all NaN
    idx = Index(dim, "bf16", 0, n)
    idx.append(X, meta)
    idx.set_doc_tags(0, doc_tags)
    Xs = oracle.round_bf16(X)
    v = valid.astype(bool)
    doc = meta["doc_idx"]
    rng = np.random.default_rng(5)
    pool = rng.choice(info["n_docs"], size=40, replace=False)          # most tiles are skipped entirely
    cases = [
        (None, v),
        (Filter().doc_pool(pool), np.isin(doc, pool) & v),
        (Filter().state_eq(synth.STATES.index("TX")), (meta["state"] == synth.STATES.index("TX")) & v),
        (Filter().doc_eq(int(doc[dup[0]])), (doc == doc[dup[0]]) & v),
        (Filter().payer_in([0xFFFE]), np.zeros(n, bool)),
    ]
    for flt, want in cases:
        s, r, c = idx.search(Q, k, flt, options=N.OPT_FORCE_MMA)
        assert idx.last_scan_kind() == "mma"
        check_all(oracle, Xs, Q, want, k, s, r, c, "bf16")
        s2, r2, c2 = idx.search(Q, k, flt, options=N.OPT_FORCE_GEMV)
        # the two scans agree with each other wherever the oracle's order is unambiguous
        assert (c == c2).all()
    s, r, c = idx.search(Q[:1], 50, options=N.OPT_FORCE_MMA)
    assert r[0].tolist() == dup[:50].tolist() or r[0, :50].tolist() == sorted(r[0, :50].tolist())
    idx.close()


def test_mma_threshold_sampling_pass(oracle):
    """Large shards first scan every 64th tile to get an admission bound per query; run it on a small
    shard (MRAG_SAMPLE_MIN_TILES=1 in a child process) incl. duplicates that tie exactly at the bound."""
    import subprocess, sys, textwrap, os
    code = textwrap.dedent('''
        import numpy as np, sys
        sys.path.insert(0, %r)
        import mrag_b200
        from mrag_b200 import synth, _native as N
        from mrag_b200.index import Index, make_meta, Filter
        from oracle import oracle
        n, dim = 150000, 128
        X, valid = synth.make_corpus(n, dim, seed=77, null_frac=1e-3)
        dup = np.arange(0, n, 64 * 64)[:40] + 3          # duplicates sitting exactly in sampled tiles
        X[dup] = X[dup[0]]; valid[dup] = 1
        Q = synth.make_queries(X, 70, seed=78)
        Q[0] = X[dup[0]]
        idx = Index(dim, "bf16", 0, n)
        idx.append(X, make_meta(n, valid=valid))
        Xs = oracle.round_bf16(X)
        for k in (10, 100, 200):
            s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA)
            for i in range(Q.shape[0]):
                oracle.check_topk(r[i], s[i], int(c[i]), oracle.all_similarities(Xs, Q[i]), valid.astype(bool), k, rtol=1e-2)
        # a filter that empties most sampled tiles: the bound may be absent for some queries
        flt = Filter().doc_pool(list(range(0, n, 997)))
        s, r, c = idx.search(Q[:5], 10, flt, options=N.OPT_FORCE_MMA)
        mask = valid.astype(bool) & np.isin(np.arange(n), np.arange(0, n, 997))
        for i in range(5):
            oracle.check_topk(r[i], s[i], int(c[i]), oracle.all_similarities(Xs, Q[i]), mask, 10, rtol=1e-2)
        print("SAMPLING-OK")
    ''') % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MRAG_SAMPLE_MIN_TILES="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "SAMPLING-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


# ---------------------------------------------------------------------------------------------
# large batches: 128-query candidate scan (fp16 queries x bf16 rows / bf16 shadow) + exact rescoring
# with certificate + exact rescan of uncertified queries.  Results must be EXACT, not approximate.
# ---------------------------------------------------------------------------------------------
def _fallbacks():
    return int(N.load().mrag_debug_fallback_count())


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,dim,nq,k", [
    (30000, 768, 130, 10), (5000, 96, 70, 32), (64, 64, 5, 10), (63, 64, 9, 1), (20000, 768, 256, 10),
    (4097, 100, 129, 7), (100000, 768, 200, 10), (300, 768, 140, 20), (30000, 768, 130, 100), (9000, 128, 70, 64),
    (20000, 256, 150, 106),
])
def test_mma128_rescore_exact(oracle, dtype, n, dim, nq, k):
    X, valid = synth.make_corpus(n, dim, seed=n + dim + 2, null_frac=3e-3)
    Q = synth.make_queries(X, nq, seed=k + 2)
    idx = Index(dim, dtype, 0, n + 5)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA128)
    assert idx.last_scan_kind() == "mma128"
    check_all(oracle, stored(oracle, X, dtype), Q, valid.astype(bool), k, s, r, c, dtype)
    assert 0 <= _fallbacks() <= nq
    idx.close()


@pytest.mark.parametrize("wide", ["1", "0"])
@pytest.mark.parametrize("n,dim,nq,k", [(300_001, 768, 300, 10), (70_000, 320, 257, 40), (129, 64, 130, 3),
                                        # 9 / 10 / 11 k-blocks: 1 / 2 / 3 query k-blocks in shared memory; 576 = a pipeline
                                        # stage (3 k-blocks) that mixes shared-memory and tensor-memory query operands
                                        (20_000, 576, 200, 10), (20_000, 640, 129, 5), (9_000, 704, 256, 10)])
def test_pair_scan_both_tile_shapes(oracle, monkeypatch, wide, n, dim, nq, k):
    """More than 128 queries: CTA pairs (cta_group::2).  MRAG_MMA256W=1 (default) = 128-row tiles, one accumulator,
    8 select warps; =0 = 64-row tiles, two accumulators.  Ragged tails: rows % 128 != 0, a last pass whose second CTA
    has no queries, a shard smaller than one tile."""
    monkeypatch.setenv("MRAG_MMA256W", wide)
    X, valid = synth.make_corpus(n, dim, seed=n % 1000 + dim, null_frac=2e-3)
    Q = synth.make_queries(X, nq, seed=k + 40)
    idx = Index(dim, "bf16", 0, n + 3)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA128)
    assert idx.last_scan_kind() == "mma128"
    check_all(oracle, stored(oracle, X, "bf16"), Q, valid.astype(bool), k, s, r, c, "bf16")
    # the exact rescan would hide a broken candidate scan: on this data at most a few certificates fail
    assert _fallbacks() <= max(2, nq // 10), _fallbacks()
    idx.close()


def test_mma128_default_dispatch(oracle):
    n, dim = 8000, 768
    X, valid = synth.make_corpus(n, dim, seed=5)
    for dtype, nq, want in (("f32", 4, "gemv"), ("f32", 5, "mma128"), ("bf16", 64, "mma"), ("bf16", 65, "mma128")):
        idx = Index(dim, dtype, 0, n)
        idx.append(X, make_meta(n, valid=valid))
        Q = synth.make_queries(X, nq, seed=6)
        s, r, c = idx.search(Q, 10)
        assert idx.last_scan_kind() == want, (dtype, nq)
        check_all(oracle, stored(oracle, X, dtype), Q, valid.astype(bool), 10, s, r, c, dtype)
        s, r, c = idx.search(Q, 107)                   # k > 106: the exact kernels
        assert idx.last_scan_kind() != "mma128"
        with pytest.raises(N.MragError):
            idx.search(Q, 107, options=N.OPT_FORCE_MMA128)
        idx.close()
    idx = Index(1536, "bf16", 0, 100)                  # dim > 768: no tensor-core path
    idx.append(np.ones((10, 1536), np.float32))
    with pytest.raises(N.MragError):
        idx.search(np.ones((1, 1536), np.float32), 5, options=N.OPT_FORCE_MMA128)
    idx.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_mma128_certificate_failure_goes_to_exact_rescan(oracle, dtype, monkeypatch):
    """A huge eps makes every certificate fail: all queries take the exact CUDA-core rescan, and the
    answer must not change.  With the real eps on data full of near-duplicates some queries fail."""
    n, dim, nq, k = 20000, 128, 70, 10
    X, valid = synth.make_corpus(n, dim, seed=9, null_frac=2e-3)
    # 300 near-copies of one row (spread 1e-4): more near-ties than k + 32 candidates can hold
    rng = np.random.default_rng(1)
    X[5000:5300] = X[77] * (1.0 + 1e-4 * rng.standard_normal((300, 1)).astype(np.float32)) \
        + 1e-4 * rng.standard_normal((300, dim)).astype(np.float32)
    Q = synth.make_queries(X, nq, seed=10)
    Q[0] = X[77] + 1e-3 * rng.standard_normal(dim).astype(np.float32)
    idx = Index(dim, dtype, 0, n)
    idx.append(X, make_meta(n, valid=valid))
    Xs = stored(oracle, X, dtype)
    s, r, c = idx.search(Q, k, options=N.OPT_FORCE_MMA128)
    check_all(oracle, Xs, Q, valid.astype(bool), k, s, r, c, dtype)
    assert _fallbacks() >= 1                           # query 0 sits in the dense cluster
    monkeypatch.setenv("MRAG_APPROX_EPS_SCALE", "1e6")
    s2, r2, c2 = idx.search(Q, k, options=N.OPT_FORCE_MMA128)
    assert _fallbacks() == nq
    check_all(oracle, Xs, Q, valid.astype(bool), k, s2, r2, c2, dtype)
    idx.close()


def test_mma128_with_filters_ties_and_nan(oracle):
    n, dim, nq, k = 30000, 256, 150, 10
    X, valid = synth.make_corpus(n, dim, seed=21, null_frac=5e-3)
    X[100:110] = 0.0                                   # zero-norm rows: NaN similarity, sort last
    X[2000:2040] = X[1999]                             # exact duplicates: ties by row
    meta, doc_tags, info = synth.make_metadata(n, seed=22, rows_per_doc=32, valid=valid)
    Q = synth.make_queries(X, nq, seed=23)
    Q[3] = 0.0                                         # zero query: every similarity NaN
    Q[4] = X[1999]
    for dtype in ("f32", "bf16"):
        idx = Index(dim, dtype, 0, n)
        idx.append(X, meta)
        idx.set_doc_tags(0, doc_tags)
        Xs = stored(oracle, X, dtype)
        for flt, m in (
            (None, valid.astype(bool)),
            (Filter().state_eq(synth.STATES.index("FL")), valid.astype(bool) & (meta["state"] == synth.STATES.index("FL"))),
            (Filter().doc_pool(np.arange(0, info["n_docs"], 37)), valid.astype(bool) & (info["doc_of_row"] % 37 == 0)),
            (Filter().doc_eq(3), valid.astype(bool) & (info["doc_of_row"] == 3)),       # fewer rows than k + 32
        ):
            s, r, c = idx.search(Q, k, flt, options=N.OPT_FORCE_MMA128)
            check_all(oracle, Xs, Q, m, k, s, r, c, dtype)
        idx.close()


# ---------------------------------------------------------------------------------------------
# fp32 shards of any width / small batches: candidates from the bf16 shadow on the CUDA cores,
# exact rescoring + certificate + exact rescan (same machinery as the 128-query path)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dim,nq,k", [(20000, 1536, 1, 10), (9000, 1536, 6, 10), (30000, 768, 3, 10), (5000, 100, 2, 100),
                                        (64, 64, 1, 5), (3000, 2000, 1, 40)])
def test_gemv_shadow_rescore_exact(oracle, monkeypatch, n, dim, nq, k):
    monkeypatch.setenv("MRAG_SHADOW_GEMV_MIN_ELEMS", "0")
    X, valid = synth.make_corpus(n, dim, seed=n + dim + 3, null_frac=3e-3)
    X[10:14] = 0.0
    Q = synth.make_queries(X, nq, seed=k + 3)
    idx = Index(dim, "f32", 0, n + 5)
    idx.append(X, make_meta(n, valid=valid))
    s, r, c = idx.search(Q, k)
    assert idx.last_scan_kind() == "gemv_shadow"
    check_all(oracle, X, Q, valid.astype(bool), k, s, r, c, "f32")
    monkeypatch.setenv("MRAG_APPROX_EPS_SCALE", "1e6")           # every certificate fails -> exact rescan
    s2, r2, c2 = idx.search(Q, k)
    assert _fallbacks() == nq
    check_all(oracle, X, Q, valid.astype(bool), k, s2, r2, c2, "f32")
    monkeypatch.setenv("MRAG_SHADOW_GEMV_MIN_ELEMS", str(1 << 60))
    idx.search(Q, k)
    assert idx.last_scan_kind() == "gemv"                          # small shards keep the exact fp32 scan
    idx.close()


def test_coalesced_concurrent_searches(oracle):
    """MRAG_OPT_COALESCE: single-query searches from many host threads are served in shared passes and every
    caller gets exactly what a lone search returns."""
    import ctypes as C
    n, dim, k = 150_000, 256, 10
    X, valid = synth.make_corpus(n, dim, seed=77)
    idx = Index(dim, "bf16", 0, n)
    idx.append(X, make_meta(n, valid=valid))
    T, per = 16, 12
    Q = synth.make_queries(X, T * per, seed=78)
    want = [idx.search(Q[i:i + 1], k) for i in range(T * per)]
    got = [None] * (T * per)
    errs = []

    def worker(t):
        try:
            for j in range(per):
                i = t * per + j
                got[i] = idx.search(Q[i:i + 1], k, options=N.OPT_COALESCE)
        except Exception as e:   # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs
    for g, w in zip(got, want):
        assert (g[1] == w[1]).all() and (g[2] == w[2]).all()
        assert np.abs(g[0] - w[0]).max() <= 2e-6
    stats = (C.c_int64 * 2)()
    N.load().mrag_debug_coalesce_stats(idx._h, stats)
    assert stats[1] == T * per and 1 <= stats[0] <= stats[1]
    # mixed k and a filtered call (never coalesced) still work side by side
    s1 = idx.search(Q[:1], 5, options=N.OPT_COALESCE)
    s2 = idx.search(Q[:1], 5, Filter().doc_eq(7), options=N.OPT_COALESCE)
    assert s1[2][0] == 5 and s2[2][0] <= 5
    idx.close()


def test_filter_mask_random_combinations():
    """K2 is integer / bitset work: 150 random conjunctions of every clause kind must reproduce, bit for bit, the
    same predicate evaluated with numpy on the codes."""
    n, dim = 40000, 16
    X, valid = synth.make_corpus(n, dim, seed=51, null_frac=5e-3)
    meta, doc_tags, info = synth.make_metadata(n, seed=52, rows_per_doc=24, valid=valid)
    idx = Index(dim, "f32", 0, n)
    idx.append(X, meta)
    idx.set_doc_tags(0, doc_tags)
    v = valid.astype(bool)
    doc = meta["doc_idx"]
    n_docs = info["n_docs"]
    tagmat = info["tagmat"]
    rng = np.random.default_rng(53)
    nP, nS, nPr, nA, nT = len(synth.PAYERS), len(synth.STATES), len(synth.PROGRAMS), len(synth.AUTHORITIES), len(synth.SOURCE_TYPES)
    for case in range(150):
        f = Filter()
        want = v.copy()
        kinds = rng.choice(8, size=int(rng.integers(1, 5)), replace=False)
        for kind in kinds:
            if kind == 0:
                codes = rng.choice(nP, size=int(rng.integers(1, 4)), replace=False)
                alt = rng.choice(nP, size=int(rng.integers(0, 3)), replace=False)
                st = int(rng.integers(0, nS))
                if len(alt):
                    f.payer_in(codes, alt, st)
                    want &= np.isin(meta["payer"], codes) | (np.isin(meta["payer"], alt) & (meta["state"] == st))
                else:
                    f.payer_in(codes)
                    want &= np.isin(meta["payer"], codes)
            elif kind == 1:
                c = int(rng.integers(0, nS)); f.state_eq(c); want &= meta["state"] == c
            elif kind == 2:
                c = int(rng.integers(0, nPr)); f.program_eq(c); want &= meta["program"] == c
            elif kind == 3:
                c = int(rng.integers(0, nA)); f.authority_eq(c); want &= meta["authority"] == c
            elif kind == 4:
                c = int(rng.integers(0, nT)); f.source_type_eq(c); want &= meta["source_type"] == c
            elif kind == 5:
                pool = rng.choice(n_docs, size=int(rng.integers(0, 400)), replace=False)
                f.doc_pool(pool); want &= np.isin(doc, pool)
            elif kind == 6:
                bits = rng.choice(64, size=int(rng.integers(1, 6)), replace=False)
                f.tag_relaxed(bits); want &= tagmat[:, bits].any(axis=1)[doc]
            else:
                sc = rng.choice(nS, size=int(rng.integers(0, 3)), replace=False)
                pc = rng.choice(nPr, size=int(rng.integers(0, 3)), replace=False)
                yc = rng.choice(nP, size=int(rng.integers(0, 3)), replace=False)
                f.tag_strict(sc, pc, yc)
                want &= np.isin(meta["state"], sc) | np.isin(meta["program"], pc) | np.isin(meta["payer"], yc)
        bits, n_pass = idx.filter_mask(f)
        got = np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (got == want).all(), (case, kinds)
        assert n_pass == int(want.sum())
    idx.close()


def test_snapshot_round_trip(oracle, tmp_path):
    """save -> load gives a table that answers every kind of call byte for byte (rows, NULL / deleted rows,
    filters, document tags, fp32 shadow rebuilt on load)."""
    from helpers import build_tables
    ot, pt, X, valid, meta, info = build_tables(oracle, 5000, 96, seed=17, dtype="f32")
    pt.delete_document(ot.document_id[100])
    pt.save(str(tmp_path / "snap"), corpus_version=42)
    pt2, ver = mrag_b200.PublishedTable.load(str(tmp_path / "snap"), device=0, capacity=6000)
    assert ver == 42 and len(pt2) == len(pt)
    rng = np.random.default_rng(5)
    Q = synth.make_queries(X, 9, seed=18)
    for flt in (None, Filter().state_eq(0), Filter().tag_relaxed([0, 3])):
        a = pt.index.search(Q, 20, flt)
        b = pt2.index.search(Q, 20, flt)
        assert all((x == y).all() or (np.isnan(x) == np.isnan(y)).all() for x, y in zip(a, b))
        assert (a[1] == b[1]).all() and (a[2] == b[2]).all()
    emb = X[int(rng.integers(0, len(X)))].tolist()
    assert mrag_b200.vector_arm(pt, emb, 10, None, None) == mrag_b200.vector_arm(pt2, emb, 10, None, None)
    # the loaded table keeps working as a table: insert + search
    pt2.insert([{"id": "new-1", "document_id": "doc-new", "source_type": "fact", "source_id": "s"}], [X[3].tolist()])
    hit = mrag_b200.B200VectorStore(table=pt2).search(X[3].tolist(), 3)
    assert any(h["id"] == "new-1" for h in hit)
    with pytest.raises(N.MragError):
        Index.load(str(tmp_path / "snap" / "table.json"))         # not a snapshot file
    pt.index.close(); pt2.index.close()
