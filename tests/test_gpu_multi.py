"""ONE process owning several row shards of one table (MultiIndex behind PublishedTable / B200VectorStore /
vector_arm): the plugin-boundary calls with the reference's signatures (vector_store.py:181-226,
corpus_search.py:1427-1438) must return exactly what the unsharded table returns -- ids, order (ties included: a row's
id is its position in the host table, so the K4 merge breaks ties like one index), scores, counts.

The shards sit on GPU 0 here (``devices=[0, 0, 0]``), which is what a 1-GPU box can run; with ``MRAG_DEVICES=0,..,7`` the
same code spreads them over the GPUs of a box (test_spread_over_all_visible_gpus runs when there is more than one).
"""
import asyncio
import threading

import numpy as np
import pytest

import mrag_b200
from mrag_b200 import synth
from mrag_b200.corpus_search import CorpusFilters, LexiconExpansion
from mrag_b200.index import Filter, make_meta
from mrag_b200.multi import MultiIndex

from helpers import build_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(oracle):
    """the same 6000-row table, unsharded and as three shards"""
    ot, one, X, valid, meta, info = build_tables(oracle, 6000, 96, seed=31, dtype="f32")
    _, three, *_ = build_tables(oracle, 6000, 96, seed=31, dtype="f32", devices=[0, 0, 0])
    yield ot, one, three, X
    one.index.close()
    three.index.close()


def test_documents_stay_on_one_shard_and_shards_balance(pair):
    _, _, three, _ = pair
    mi = three.index
    assert isinstance(mi, MultiIndex)
    sizes = mi.shard_sizes()
    assert sum(sizes) == len(three) == 6000 and max(sizes) - min(sizes) < 200
    sh = mi.pos_shard[:6000]
    docs = three.row_doc[:6000]
    for d in np.unique(docs)[::7]:
        assert len(set(sh[docs == d].tolist())) == 1, "a document's chunks must share a shard"


def test_store_search_equals_unsharded(pair):
    ot, one, three, X = pair
    s1, s3 = mrag_b200.B200VectorStore(table=one), mrag_b200.B200VectorStore(table=three)
    doc = ot.document_id[1234]
    rng = np.random.default_rng(3)
    for kw in (dict(k=10), dict(k=1), dict(k=100), dict(k=10, document_id=doc), dict(k=25, filters={"payer": "Sunshine Health"}),
               dict(k=10, filters={"state": "FL", "authority_level": "payer_policy"}), dict(k=10, filters={"payer": "nobody"}),
               dict(k=300)):
        for _ in range(3):
            q = (X[int(rng.integers(0, 6000))] + 0.05 * rng.standard_normal(96)).tolist()
            assert s3.search(q, **kw) == s1.search(q, **kw)
    # an exact duplicate cluster: ties must come back in the same (host position) order
    dup = X[40].tolist()
    assert s3.search(dup, 50) == s1.search(dup, 50)

    async def both():
        return await asyncio.gather(s3.asearch(dup, 7), s1.asearch(dup, 7))
    a, b = asyncio.run(both())
    assert a == b and len(a) == 7


def test_vector_arm_equals_unsharded(pair):
    ot, one, three, X = pair
    rng = np.random.default_rng(4)
    docs = sorted(set(ot.document_id))
    pool = [docs[int(j)] for j in rng.choice(len(docs), 30, replace=False)]
    cases = [
        dict(k=10, filters=None, include_document_ids=None),
        dict(k=20, filters=CorpusFilters(payer="Sunshine Health"), include_document_ids=None),          # FL-MCO union
        dict(k=10, filters=CorpusFilters(state="FL", program="Medicaid"), include_document_ids=pool),
        dict(k=10, filters=None, include_document_ids=pool, over_fetch_factor=8, min_similarity=0.05),
        dict(k=10, filters=None, include_document_ids=None, tag_mode="auto",
             expansion=LexiconExpansion(jurisdiction_tags=["j:state.zz"], domain_tags=["d:topic_000.leaf"])),   # strict -> relaxed retry
        dict(k=10, filters=None, include_document_ids=None, tag_mode="relaxed",
             expansion=LexiconExpansion(domain_tags=["d:topic_006.leaf"], process_tags=["p:topic_001.leaf"])),
    ]
    for kw in cases:
        kw = dict(kw)
        k, filters, pool_ids = kw.pop("k"), kw.pop("filters"), kw.pop("include_document_ids")
        q = (X[int(rng.integers(0, 6000))] + 0.1 * rng.standard_normal(96)).tolist()
        got = mrag_b200.vector_arm(three, q, k, filters, pool_ids, **kw)
        want = mrag_b200.vector_arm(one, q, k, filters, pool_ids, **kw)
        assert got == want and (len(want) > 0 or kw.get("tag_mode") == "auto" or True)


def test_oracle_parity_through_the_sharded_table(oracle, pair):
    ot, _, three, X = pair
    rng = np.random.default_rng(6)
    for i in range(4):
        q = (X[int(rng.integers(0, 6000))] + 0.1 * rng.standard_normal(96)).tolist()
        want = oracle.vector_arm(ot, q, 15, None, None)
        got = mrag_b200.vector_arm(three, q, 15, None, None)
        assert [g["id"] for g in got] == [w["id"] for w in want]
        assert all(g["similarity"] == pytest.approx(w["similarity"], rel=1e-4, abs=1e-6) for g, w in zip(got, want))


def test_delete_republish_and_incremental_appends(oracle):
    ot, pt, X, valid, meta, info = build_tables(oracle, 1500, 64, seed=41, dtype="f32", devices=[0, 0])
    _, ref, *_ = build_tables(oracle, 1500, 64, seed=41, dtype="f32")
    doc = ot.document_id[400]
    for t in (pt, ref):
        n = t.delete_document(doc)
        assert n == sum(1 for d in ot.document_id if d == doc)
        # re-publish the document in worker-sized batches of 50 rows that split it across appends
        rows = [{"id": f"new-{i}", "document_id": doc, "source_type": "fact", "source_id": f"s{i}"} for i in range(60)]
        for lo in range(0, 60, 25):
            t.insert(rows[lo:lo + 25], [X[(400 + i) % 1500].tolist() for i in range(lo, min(lo + 25, 60))])
    s2, s1 = mrag_b200.B200VectorStore(table=pt), mrag_b200.B200VectorStore(table=ref)
    for j in (400, 3, 900):
        assert s2.search(X[j].tolist(), 20) == s1.search(X[j].tolist(), 20)
        assert s2.search(X[j].tolist(), 20, document_id=doc) == s1.search(X[j].tolist(), 20, document_id=doc)
    assert len({pt.index.pos_shard[r] for r in range(1500, 1560)}) == 1      # the re-published document stays together
    pt.index.close(); ref.index.close()


def test_snapshot_round_trip_sharded(oracle, tmp_path):
    ot, pt, X, *_ = build_tables(oracle, 2000, 64, seed=43, dtype="bf16", devices=[0, 0, 0])
    pt.save(str(tmp_path / "snap"), corpus_version=7)
    pt2, ver = mrag_b200.PublishedTable.load(str(tmp_path / "snap"), devices=[0, 0, 0], capacity=4000)
    assert ver == 7 and len(pt2) == 2000 and isinstance(pt2.index, MultiIndex)
    s1, s2 = mrag_b200.B200VectorStore(table=pt), mrag_b200.B200VectorStore(table=pt2)
    for j in (5, 700, 1999):
        assert s1.search(X[j].tolist(), 12) == s2.search(X[j].tolist(), 12)
    pt2.insert([{"id": "late", "document_id": "doc-late"}], [X[9].tolist()])
    assert any(h["id"] == "late" for h in s2.search(X[9].tolist(), 3))
    with pytest.raises(ValueError):
        mrag_b200.PublishedTable.load(str(tmp_path / "snap"), devices=[0, 0])      # shard count must match
    pt.index.close(); pt2.index.close()


def test_searches_concurrent_with_inserts_always_hydrate(oracle):
    """a search that lands between the device append and the host append used to index past the host
    columns.  Host columns are written first now; hammer both tables with searches while a writer inserts."""
    for devices in (None, [0, 0]):
        pt = mrag_b200.PublishedTable(48, "f32", 0, 40000, devices=devices)
        rng = np.random.default_rng(8)
        X = rng.standard_normal((30000, 48)).astype(np.float32)
        store = mrag_b200.B200VectorStore(table=pt)
        pt.insert([{"id": "seed", "document_id": "d0"}], [X[0].tolist()])
        stop, errors = threading.Event(), []

        def reader():
            q = X[1].tolist()
            while not stop.is_set():
                try:
                    for h in store.search(q, 50):
                        assert h["id"] is not None and h["document_id"].startswith("d")
                    assert isinstance(mrag_b200.vector_arm(pt, q, 10, None, None), list)
                except Exception as e:          # noqa: BLE001
                    errors.append(repr(e))
                    return
        th = [threading.Thread(target=reader) for _ in range(3)]
        [t.start() for t in th]
        for lo in range(1, 30000, 500):
            pt.insert([{"id": f"r{i}", "document_id": f"d{i // 40}"} for i in range(lo, min(lo + 500, 30000))],
                      [X[i].tolist() for i in range(lo, min(lo + 500, 30000))])
        stop.set()
        [t.join() for t in th]
        assert not errors, errors[:3]
        assert len(pt) == 30000
        pt.index.close()


def test_bulk_column_insert_matches_row_insert(oracle):
    n, dim = 3000, 64
    X, valid = synth.make_corpus(n, dim, seed=51)
    ids = [f"id-{i}" for i in range(n)]
    docs = [f"doc-{i // 37}" for i in range(n)]
    cols = {"id": ids, "document_id": docs, "document_payer": ["Aetna" if i % 3 else None for i in range(n)],
            "text": [f"body {i}" for i in range(n)], "page_number": list(range(n)), "chunk_d_tags": [{"a.b": 1} if i % 9 == 0 else None for i in range(n)]}
    a = mrag_b200.PublishedTable(dim, "f32", 0, n)
    a.insert_columns(X, cols, valid)
    b = mrag_b200.PublishedTable(dim, "f32", 0, n)
    rows = [{k: v[i] for k, v in cols.items()} for i in range(n)]
    for lo in range(0, n, 1000):
        b.insert(rows[lo:lo + 1000], [X[i].tolist() if valid[i] else None for i in range(lo, lo + 1000)])
    q = X[77].tolist()
    assert mrag_b200.vector_arm(a, q, 20, CorpusFilters(payer="Aetna"), None) == mrag_b200.vector_arm(b, q, 20, CorpusFilters(payer="Aetna"), None)
    got = mrag_b200.vector_arm(a, X[9].tolist(), 1, None, None)[0]
    assert got["id"] == "id-9" and got["text"] == "body 9" and got["page_number"] == 9 and got["chunk_d_tags"] == {"a.b": 1} and got["payer"] is None
    a.index.close(); b.index.close()


def test_chroma_shaped_store(oracle):
    """a3: ChromaVectorStore.search (vector_store.py:77-99) -- cosine DISTANCE under `distance`, where = document_id only,
    metadata stringified on add"""
    n, dim = 800, 32
    X, valid = synth.make_corpus(n, dim, seed=61, null_frac=0.0, zero_norm_rows=0)
    store = mrag_b200.B200ChromaVectorStore(collection_name="chunk_embeddings", dim=dim, capacity=n)
    meta = [{"document_id": f"doc-{i // 20}", "source_type": None if i % 5 == 0 else "fact", "source_id": i} for i in range(n)]
    for lo in range(0, n, 50):                                   # the worker's batches (embedding_worker.py:256)
        store.add([f"c{i}" for i in range(lo, lo + 50)], [X[i].tolist() for i in range(lo, lo + 50)], meta[lo:lo + 50])
    q = X[123] + 0.05 * np.random.default_rng(1).standard_normal(dim).astype(np.float32)
    got = store.search(q.tolist(), 10)
    sims = 1.0 - oracle.cosine_distance_c(X, q.astype(np.float32))
    order = np.argsort(-np.nan_to_num(sims, nan=-9.0), kind="stable")[:10]
    assert [g["id"] for g in got] == [f"c{i}" for i in order]
    assert set(got[0]) == {"id", "document_id", "source_type", "source_id", "distance"}
    for g, i in zip(got, order):
        assert g["distance"] == pytest.approx(1.0 - sims[i], abs=1e-5) and 0.0 <= g["distance"] <= 2.0
        assert g["document_id"] == f"doc-{i // 20}" and g["source_id"] == str(i)
        assert g["source_type"] == ("None" if i % 5 == 0 else "fact")          # str(None), as Chroma's add stringifies
    only = store.search(q.tolist(), 100, document_id="doc-6")
    assert len(only) == 20 and all(g["document_id"] == "doc-6" for g in only)
    assert [g["distance"] for g in only] == sorted(g["distance"] for g in only)
    store.delete_by_document("doc-6")
    assert store.search(q.tolist(), 100, document_id="doc-6") == []
    assert store.search(q.tolist(), 5, document_id="no-such-doc") == []
    store.table.index.close()


def test_spread_over_all_visible_gpus(oracle):
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("one GPU visible: the multi-device spread is covered by devices=[0, 0, 0] above")
    ot, pt, X, *_ = build_tables(oracle, 4000, 64, seed=71, dtype="bf16", devices=list(range(g)))
    _, ref, *_ = build_tables(oracle, 4000, 64, seed=71, dtype="bf16")
    s2, s1 = mrag_b200.B200VectorStore(table=pt), mrag_b200.B200VectorStore(table=ref)
    for j in (1, 2222, 3999):
        assert s2.search(X[j].tolist(), 30) == s1.search(X[j].tolist(), 30)
    pt.index.close(); ref.index.close()
